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
