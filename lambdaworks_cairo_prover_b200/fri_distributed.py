"""fri_commit_phase / fri_query_phase / grinding for ONE proof on the GPUs of one box.

Row-block distribution, one process per GPU: rank r holds rows [r*size/G, (r+1)*size/G) of every sharded layer.

  commit   a layer is a one-column table, so its tree is the row-block tree of distributed.py: every rank hashes its
           rows and builds that subtree, the G subtree roots are all-gathered, the top log2(G) levels are hashed on
           every rank (FriMerkleTree == BatchedMerkleTree over one column, src/starks/config.rs:10-20).
  fold     out[i] = (v + s)/2 + zeta (v - s)/(2 h_k w^i) with v = layer[i], s = layer[i + size/2] (the verifier's
           formula, verifier.rs:511-512; same values as fold_polynomial + a fresh FFT, fri/mod.rs:43-54).  The
           partner of row i lives G/2 ranks away: every rank sends the two halves of its block to the two ranks
           that fold them -- one pairwise exchange of half a layer per fold (ncclSend/ncclRecv over NVLink).
  collapse once a layer is small (<= 2^19 evaluations: its folds are latency, not throughput) it is all-gathered and rank 0 finishes the
           phase with the single-GPU path (device-side transcript chain + tail kernel); the other ranks replay the
           transcript from the roots rank 0 broadcasts.
  queries  the owner of a row serves its value and the subtree part of its path; the top levels are replicated.
  grinding rank r searches batches r, r+G, .. of every 2^32-nonce window; MIN all-reduce = the reference's nonce.

Every rank runs the same Fiat-Shamir transcript, so no challenge is ever sent.  The compute steps go through a backend
object (GpuCairoBackend on GPUs; the tests substitute a CPU double so that this file also runs under gloo).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import distributed as D

U64_MAX = (1 << 64) - 1


class ShardedFri:
    """What one rank holds after the sharded commit phase."""

    def __init__(self, sharded_layers, tail, tail_first, roots, last_value, domain_size):
        self.sharded_layers = sharded_layers       # ShardedCommit per layer k < tail_first
        self.tail = tail                           # rank 0: backend handle for layers tail_first..L-1 (None elsewhere)
        self.tail_first = tail_first
        self.roots = roots                         # all L layer roots (bytes)
        self.last_value = last_value               # LW element, uint64[4]
        self.domain_size = domain_size

    def free(self, be):
        for sc in self.sharded_layers:
            sc.free()
        if self.tail is not None:
            be.release_fri(self.tail)
        self.sharded_layers, self.tail = [], None


def _grank(group, r):
    return r if group is None else dist.get_global_rank(group, r)


def fold_exchange(block, rank, world, group, be):
    """The pairwise exchange before a fold.  block: this rank's rows of the layer, [B, 4].  Returns (v, s): the
    B/2 values layer[i] and layer[i + size/2] for this rank's rows i of the NEXT layer."""
    b = block.shape[0]
    half = b // 2
    if world == 1:
        return block[:half], block[half:]
    lo, hi = block[:half], block[half:]
    base = 2 * (rank % (world // 2))               # the two ranks that fold my rows
    am_v = rank < world // 2                       # first half of the layer: my rows are the `v` operands
    v = be.new_tensor((half, 4))
    s = be.new_tensor((half, 4))
    ops = []
    for dst, piece in ((base, lo), (base + 1, hi)):
        if dst == rank:
            (v if am_v else s).copy_(piece)
        else:
            ops.append(dist.P2POp(dist.isend, piece.contiguous(), _grank(group, dst), group))
    src_v, src_s = rank // 2, rank // 2 + world // 2
    if src_v != rank:
        ops.append(dist.P2POp(dist.irecv, v, _grank(group, src_v), group))
    if src_s != rank:
        ops.append(dist.P2POp(dist.irecv, s, _grank(group, src_s), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return v, s


def commit_row_block(block_cols, n_rows_total, t, be, group, shards=None):
    """batch_commit (prover.rs:96-104) of a table whose rows are spread over the ranks in equal contiguous blocks:
    block_cols [c, rows, 4] = all columns of this rank's rows.  Appends the root to the transcript."""
    world = dist.get_world_size(group)
    block, roots = D.gather_subtree_roots(be, block_cols, group)
    top = D.build_top([roots[32 * g:32 * g + 32] for g in range(world)], be.keccak)
    t.append(top[0])
    sc = D.ShardedCommit(be, group, None, block, top, n_rows_total, block_cols.shape[0], shards)
    sc.block_tensor = block_cols
    return sc


def fri_commit_phase_sharded(p0_block, domain_size, number_layers, t, coset_offset, be, group=None, collapse_log=None):
    """fri_commit_phase (src/starks/fri/mod.rs:20-72) with layer 0 given as this rank's block of the evaluations on
    the LDE coset (p0_block [domain_size/G, 4], internal element format on GPUs).  Returns a ShardedFri; every rank's
    transcript ends in the same state (last value appended)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if collapse_log is None:
        # a layer of 2^19 evaluations costs one GPU ~0.2 ms (fold + leaves + tree); sharding it saves less than the
        # gather of the subtree roots and the fold exchange cost in latency
        collapse_log = 19
    sharded = []
    roots = []
    block = p0_block
    size = domain_size
    k = 0
    # layers stay sharded while they are large; the last committed layer always belongs to the tail
    while world > 1 and k < number_layers - 1 and size > (1 << collapse_log) and (size // world) >= 2:
        sc = commit_row_block(block.view(1, block.shape[0], 4), size, t, be, group)
        sharded.append(sc)
        roots.append(sc.root)
        zeta = be.to_field(t)                                                    # fri/mod.rs:41
        v, s = fold_exchange(block, rank, world, group, be)
        nxt = be.new_tensor((block.shape[0] // 2, 4))
        be.fri_fold_rows(v, s, rank * (block.shape[0] // 2), size, domain_size, k, zeta, coset_offset, nxt)
        block = nxt
        size //= 2
        k += 1
    # ---- collapse: layer k in full on rank 0 (all-gathered: the block sizes are equal), single-GPU path from there
    if world > 1:
        full = be.new_tensor((size, 4))
        dist.all_gather_into_tensor(full.view(-1), block.reshape(-1).contiguous(), group=group)
    else:
        full = block
    n_tail = number_layers - k
    tail = None
    payload = np.zeros((n_tail + 1, 32), dtype=np.uint8)                          # tail roots + last value (LW bytes)
    if rank == 0:
        tail, last, tail_roots = be.fri_continue(full, n_tail, t, coset_offset, k, domain_size)
        payload[:n_tail] = np.asarray(tail_roots, dtype=np.uint8).reshape(n_tail, 32)
        payload[n_tail] = np.asarray(last, dtype=np.uint64).view(np.uint8)
    if world > 1:
        pt = torch.from_numpy(payload).to(be.device)
        dist.broadcast(pt, src=_grank(group, 0), group=group)
        payload = pt.cpu().numpy()
    last = payload[n_tail].view(np.uint64).copy()
    tail_roots = [payload[j].tobytes() for j in range(n_tail)]
    if rank != 0:
        # replay what rank 0's transcript went through (fri/mod.rs:37,41,54,58,69)
        for j, r in enumerate(tail_roots):
            if j > 0:
                be.to_field(t)
            t.append(r)
        be.to_field(t)
        t.append(be.to_bytes_be(last))
    roots.extend(tail_roots)
    return ShardedFri(sharded, tail, k, roots, last, domain_size)


def fri_query_sharded(fri, iotas, number_layers, be, group=None):
    """fri_query_phase's reads (src/starks/fri/mod.rs:74-127).  Rank 0 gets (ev, evs, pa, pas) shaped like
    s252_fri_query's outputs ([Q, L, 4] values, [Q, L, depth, 32] paths padded with zeros); other ranks get None."""
    rank = dist.get_rank(group)
    q = len(iotas)
    depth = fri.domain_size.bit_length() - 1
    ev = np.zeros((q, number_layers, 4), dtype=np.uint64)
    evs = np.zeros_like(ev)
    pa = np.zeros((q, number_layers, depth, 32), dtype=np.uint8)
    pas = np.zeros_like(pa)
    k0 = fri.tail_first
    if k0 > 0:
        index_lists = []
        for k in range(k0):
            size = fri.domain_size >> k
            index_lists.append([i % size for i in iotas] + [(i + size // 2) % size for i in iotas])
        opened = D.open_many_packed(fri.sharded_layers, index_lists, group)
        if rank == 0:
            for k, (rows, paths) in enumerate(opened):
                ev[:, k], evs[:, k] = rows[:q, 0], rows[q:, 0]
                pa[:, k, :depth - k], pas[:, k, :depth - k] = paths[:q], paths[q:]
    if rank != 0:
        return None
    if number_layers > k0:
        tev, tevs, tpa, tpas = be.fri_query(fri.tail, iotas, number_layers - k0, depth - k0)
        ev[:, k0:], evs[:, k0:] = tev, tevs
        pa[:, k0:, :depth - k0], pas[:, k0:, :depth - k0] = tpa, tpas
    return ev, evs, pa, pas


def generate_nonce_with_grinding_sharded(challenge, grinding_factor, be, group=None):
    """generate_nonce_with_grinding (src/starks/grinding.rs:40-48) with the search split over the ranks; every rank
    returns the same nonce: the global minimum."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    # a rank does not see the others' hits while its kernel runs: windows of about twice the expected position of the
    # first hit keep every rank from searching until its OWN first hit (at least one 2^18-nonce batch per rank)
    window_log = min(32, max(18 + (world - 1).bit_length(), grinding_factor + 1))
    base = 0
    while base < U64_MAX:
        found = be.grind_round(challenge, grinding_factor, base, rank, world, window_log)
        best = torch.tensor([found if found < (1 << 63) else (1 << 63) - 1], dtype=torch.int64, device=be.device)
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
        best = int(best.item())
        if best != (1 << 63) - 1:
            return best
        base += 1 << window_log
    raise RuntimeError("nonce not found")
