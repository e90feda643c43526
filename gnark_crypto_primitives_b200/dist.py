"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the only exchange step.

Hashes, proofs and ballots are independent, so each rank works on a contiguous index slice and no data-path
collective is needed.  The single exchange is the ElGamal tally: every rank reduces its slice to
`n_fields` partial ciphertexts (128 bytes each, canonical affine — the ordinary wire format), the partials are
all-gathered as bytes (NCCL over NVLink on GPUs, gloo in the CPU tests), and every rank tallies the gathered
(world, n_fields) array.  Edwards addition is exact and associative, so the result is bit-identical at any world size.

The reduction functions are passed in, so the same plumbing is exercised on CPU (gloo, tests/test_dist_gloo.py) with
the oracle standing in for the GPU engine; in production they are Engine.elgamal_tally_dev.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items owned by `rank`; slices differ by at most one item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return n * rank // world, n * (rank + 1) // world


def allgather_partials(partial: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """(n_fields, 4, 32) uint8 partial ciphertexts -> (world, n_fields, 4, 32), rank order, as raw bytes."""
    if partial.dtype != torch.uint8:
        raise TypeError("partials travel as bytes")
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial.unsqueeze(0).clone()
    world = dist.get_world_size(group)
    flat = partial.contiguous().view(-1)
    gathered = torch.empty(world * flat.numel(), dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(gathered, flat, group=group)
    return gathered.view((world,) + tuple(partial.shape))


def sharded_tally(local_ct: torch.Tensor, n_fields: int,
                  tally_fn: Callable[[torch.Tensor, int, int], torch.Tensor],
                  group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Tally ballots that are sharded over the ranks.

    local_ct: this rank's (n_local_ballots, n_fields, 4, 32) uint8 ciphertexts (may be empty).
    tally_fn(ct, n_ballots, n_fields) -> (n_fields, 4, 32) uint8: the per-device reduction.
    Returns the global tally on every rank.
    """
    n_local = int(local_ct.shape[0])
    partial = tally_fn(local_ct, n_local, n_fields)
    gathered = allgather_partials(partial, group)
    return tally_fn(gathered, int(gathered.shape[0]), n_fields)


def engine_tally_fn(engine, stream=None):
    """tally_fn backed by Engine.elgamal_tally_dev on CUDA tensors.  Raises if a field reports a status (one 8-byte
    read-back per call): a shard with a non-canonical ciphertext would otherwise put zeroed partials into the all-gather
    and every rank would fold them into a wrong global tally without noticing."""
    def fn(ct, n_ballots, n_fields):
        out = torch.empty((n_fields, 4, 32), dtype=torch.uint8, device=ct.device)
        status = torch.empty(n_fields, dtype=torch.uint8, device=ct.device)
        engine.elgamal_tally_dev(ct.contiguous(), n_ballots, n_fields, out, status,
                                 stream=stream if stream is not None else torch.cuda.current_stream())
        bad = status.cpu()
        if bool(bad.any()):
            raise RuntimeError(f"tally status per field: {bad.tolist()} (1 = non-canonical ciphertext, 5 = zero denominator)")
        return out
    return fn
