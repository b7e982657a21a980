"""MerkleTree<B> with the Keccak back-ends the reference selects (src/starks/config.rs:10-20):
BatchedMerkleTree (leaf = Keccak256 of a whole row) and FriMerkleTree (leaf = one element).
The tree lives on the GPU; `root` and authentication paths are read back on demand."""
import ctypes as C

import numpy as np

from . import _native as N


class Proof:
    """lambdaworks_crypto::merkle_tree::proof::Proof<Commitment>: merkle_path leaf -> root."""

    def __init__(self, merkle_path):
        self.merkle_path = [bytes(x) for x in merkle_path]


class DeviceCommit:
    """Owner of an s252_commit handle: LDE columns (+ coefficients) + tree, resident in HBM."""

    def __init__(self, ctx, handle, root):
        self.ctx, self.handle, self.root = ctx, handle, bytes(root)
        L = N.lib()
        self.n_cols = L.s252_commit_n_cols(handle)
        self.n_rows = L.s252_commit_n_rows(handle)
        self.n_coeffs = L.s252_commit_n_coeffs(handle)
        ctx.adopt(self)

    def free(self):
        if getattr(self, "handle", None):
            N.lib().s252_commit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def get_proof_by_pos(self, pos):
        """MerkleTree::get_proof_by_pos -> Option<Proof> (None when out of range)."""
        if pos < 0 or pos >= self.n_rows:
            return None
        depth = self.n_rows.bit_length() - 1
        idx = np.array([pos], dtype=np.uint64)
        paths = np.empty((max(depth, 1), 32), dtype=np.uint8)
        self.ctx.check(N.lib().s252_commit_open(self.handle, N.ptr(idx), 1, None, N.ptr(paths)))
        return Proof([paths[k].tobytes() for k in range(depth)])

    def open(self, indices):
        """Rows and authentication paths for many positions at once (open_deep_composition_poly,
        src/starks/prover.rs:484-529).  Returns (rows[q, n_cols, 4], paths[q, depth, 32])."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        depth = self.n_rows.bit_length() - 1
        rows = np.empty((len(idx), self.n_cols, 4), dtype=np.uint64)
        paths = np.empty((len(idx), max(depth, 1), 32), dtype=np.uint8)
        self.ctx.check(N.lib().s252_commit_open(self.handle, N.ptr(idx), len(idx), N.ptr(rows), N.ptr(paths)))
        return rows, paths[:, :depth]

    def lde_column(self, col, first=0, count=None):
        count = self.n_rows - first if count is None else count
        out = np.empty((count, 4), dtype=np.uint64)
        self.ctx.check(N.lib().s252_commit_read_lde(self.handle, col, first, count, N.ptr(out)))
        return out

    def coefficients(self, col):
        out = np.empty((self.n_coeffs, 4), dtype=np.uint64)
        self.ctx.check(N.lib().s252_commit_read_coeffs(self.handle, col, N.ptr(out)))
        return out

    def nodes(self, first=0, count=None):
        total = 2 * self.n_rows - 1
        count = total - first if count is None else count
        out = np.empty((count, 32), dtype=np.uint8)
        self.ctx.check(N.lib().s252_commit_read_nodes(self.handle, first, count, N.ptr(out)))
        return out


class BatchedMerkleTree(DeviceCommit):
    @staticmethod
    def build(rows, ctx=None):
        """BatchedMerkleTree::build(&rows) (src/starks/prover.rs:101): rows[n_rows][n_cols]."""
        ctx = ctx or N.default_context()
        rows = N.fe_array(np.asarray(rows, dtype=np.uint64))
        if rows.ndim == 2:
            rows = rows.reshape(rows.shape[0], 1, 4)
        h = C.c_void_p()
        root = np.empty(32, dtype=np.uint8)
        ctx.check(N.lib().s252_merkle_build(ctx.handle, N.ptr(rows), rows.shape[0], rows.shape[1], N.HOST, C.byref(h),
                                            N.ptr(root)))
        return BatchedMerkleTree(ctx, h, root.tobytes())


class FriMerkleTree(BatchedMerkleTree):
    @staticmethod
    def build(evaluation, ctx=None):
        """FriMerkleTree::build(&evaluation) (src/starks/fri/fri_commitment.rs:39)."""
        ev = np.asarray(evaluation, dtype=np.uint64).reshape(-1, 1, 4)
        t = BatchedMerkleTree.build(ev, ctx)
        t.__class__ = FriMerkleTree
        return t


def batch_commit(vectors, ctx=None):
    """src/starks/prover.rs:96-104 -> (tree, commitment)"""
    tree = BatchedMerkleTree.build(vectors, ctx)
    return tree, tree.root
