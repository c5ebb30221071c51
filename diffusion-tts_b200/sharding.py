"""Host-side mirror of the packed argmax key (csrc/sampler.cuh: argmax_first_kernel).

key = (orderable_i32(score) << 32) | (0xFFFFFFFF - global_index), a signed int64: the maximum
picks the best score and, among equal scores, the lowest global candidate index -- the rule of
torch.argmax (edm/main.py:842).  Shards combine with all_reduce(MAX) on int64."""
import torch


def pack_key(scores: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    s = scores.to(torch.float32).clone()
    s[s != s] = float('-inf')
    bits = s.contiguous().view(torch.int32).to(torch.int64)
    ordered = torch.where(bits < 0, bits ^ 0x7FFFFFFF, bits)
    return (ordered << 32) | (0xFFFFFFFF - index.to(torch.int64))


def unpack_index(key: torch.Tensor) -> torch.Tensor:
    return 0xFFFFFFFF - (key & 0xFFFFFFFF)


def exchange_winner(rows: torch.Tensor, idx: torch.Tensor, lo: int, hi: int, group=None, gather=None) -> torch.Tensor:
    """rows [hi-lo, b, ...] = this rank's slice of per-candidate tensors (candidates, committed states); idx [b] = the global
    winner per image.  The owner of each winner contributes the row, every other rank zeros, to an all_reduce(SUM)
    (x + 0 is exact): afterwards every rank holds the winners, and only b rows crossed the interconnect.
    `gather(rows, local_idx)` picks rows[local_idx[j], j] (the CUDA path passes ops.gather_rows)."""
    import torch.distributed as dist
    n_local = hi - lo
    owned = (idx >= lo) & (idx < hi)
    local = (idx - lo).clamp(0, n_local - 1).contiguous()
    if gather is None:
        gather = lambda r, li: r[li, torch.arange(r.shape[1], device=r.device)]
    out = gather(rows, local)
    # masked select, not a multiplication: a diverged (Inf/NaN) losing candidate on a non-owner must not poison the sum
    out = torch.where(owned.view(-1, *([1] * (out.dim() - 1))), out, torch.zeros_like(out))
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
