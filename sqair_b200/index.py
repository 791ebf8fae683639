"""Indexing helpers of sqair/index.py that remain meaningful at the host level.  In the fused path
IWAE tiling is virtual (particles of a sequence share its frame) and slot compaction / object ids run
inside the kernel; these torch versions exist for callers of the reference API and for tests."""
import torch


def tile_input_for_iwae(tensor, iw_samples, with_time=False):
    """index.py:106-129: repeat along the batch axis so that tiled samples are contiguous (row = b*K + k)."""
    return tensor.repeat_interleave(iw_samples, dim=1 if with_time else 0)


def select_present(x, presence, batch_size=None, name='select_present'):
    """index.py:132-165: per-row stable partition of x [B,K,d] by the binary presence [B,K], present first."""
    order = torch.argsort(1 - presence.to(torch.int64), dim=1, stable=True)
    return x.gather(1, order.reshape(order.shape + (1,) * (x.dim() - 2)).expand_as(x))


def compute_object_ids(last_used_id, prev_ids, propagated_pres, discovery_pres):
    """index.py:198-221."""
    prop_ids = prev_ids * propagated_pres - (1 - propagated_pres)
    inc = torch.cumsum(discovery_pres, 1)
    disc_ids = inc + last_used_id[:, None]
    last_used_id = last_used_id + inc[:, -1]
    disc_ids = disc_ids * discovery_pres - (1 - discovery_pres)
    return last_used_id, torch.cat([prop_ids, disc_ids], 1)


def gather_axis(tensor, idx, axis=-1):
    return tensor.index_select(axis, idx)
