"""B200-native LDE + commitment path of the lambdaworks Cairo prover.

Host-side mirror of the reference's interface for this path (same names and argument meaning),
over the C ABI in include/stark252_b200.h; the kernels are hand-written CUDA for sm_100a
(csrc/).  There is no CPU fallback.
"""
from ._native import Context, FFTError, Stark252Error, default_context, library_path  # noqa: F401
from .fri import FriDecommitment, FriLayer, fri_commit_phase, fri_open, fri_query_phase  # noqa: F401
from .grinding import generate_nonce_with_grinding  # noqa: F401
from .merkle import BatchedMerkleTree, DeviceCommit, FriMerkleTree, Proof, batch_commit  # noqa: F401
from .options import ProofOptions  # noqa: F401
from .polynomial import Polynomial, evaluate_polynomial_on_lde_domain  # noqa: F401
from .prover import (Domain, TraceTable, evaluate_at, fri_commit_phase_deep, get_trace_evaluations,  # noqa: F401
                     interpolate_and_commit, lde_and_commit)
from .transcript import DefaultTranscript, batch_sample_challenges, transcript_to_field, transcript_to_usize  # noqa: F401
