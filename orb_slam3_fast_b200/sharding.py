"""Frame sharding across the GPUs of one box: frames are independent — the extractor keeps no state
between frames (src/ORBextractor.cc:1116-1118 overwrites mvImagePyramid every call) — so the n frames are cut
into `world` contiguous blocks whose sizes differ by at most one, the first n % world ranks taking the extra frame
(shard_range below; the C entry points orbx_extract_batch_multi / orbm_stereo_track_frames_batch_multi cut the same
way); the two eyes of a stereo pair are one unit and never split. There is no data-path
collective. The only communication is the gather of fixed-size result slabs (count + cap x 28 B keypoints + cap x 32 B
descriptors per frame) to rank 0, done with torch.distributed on whatever backend the process group has (NCCL over
NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

from .lib import KP_DTYPE


def shard_range(n_units, world, rank):
    """Contiguous block of units [start, stop) of `rank`; sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_units, world):
    return [shard_range(n_units, world, r)[1] - shard_range(n_units, world, r)[0] for r in range(world)]


class FrameSharder:
    """Splits a batch of frames (or stereo pairs) over the ranks of the default process group and gathers results."""

    def __init__(self, device=None):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device = device if device is not None else torch.device("cpu")

    def my_range(self, n_units):
        return shard_range(n_units, self.world, self.rank)

    def _gather_padded(self, local, n_units):
        """local: tensor [my_units, ...] -> on every rank the tensor [n_units, ...] in global unit order."""
        if self.world == 1:
            return local
        sizes = shard_sizes(n_units, self.world)
        m = max(sizes)
        pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        out = torch.empty((self.world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad)
        parts = [out[r * m:r * m + sizes[r]] for r in range(self.world)]
        return torch.cat(parts, dim=0)

    def gather_extract_results(self, n_units, n_out, mono, kps, desc):
        """Per-rank outputs of ORBextractor.extract_batch (numpy or torch, [my_units, cap...]) -> global arrays
        (numpy) on every rank: n_out[n_units], mono[n_units], kps[n_units, cap] (KP_DTYPE), desc[n_units, cap, 32]."""
        def t(a, dtype=None):
            if isinstance(a, np.ndarray):
                if a.dtype == KP_DTYPE:
                    a = a.view(np.int32).reshape(a.shape + (7,))
                a = torch.from_numpy(np.ascontiguousarray(a))
            return a.to(self.device)
        g_n = self._gather_padded(t(n_out), n_units).cpu().numpy()
        g_m = self._gather_padded(t(mono), n_units).cpu().numpy()
        g_k = self._gather_padded(t(kps), n_units).cpu().numpy()
        g_d = self._gather_padded(t(desc), n_units).cpu().numpy()
        g_k = np.ascontiguousarray(g_k).view(KP_DTYPE).reshape(g_k.shape[:2])
        return g_n, g_m, g_k, g_d

    def gather_knn2(self, n_queries, idx1, d1, idx2, d2):
        """BASELINE configs[4] on several GPUs: the QUERY rows are sharded, every rank holds the whole
        train set (100 k x 32 B = 3.2 MB), so the 2-NN of a query never needs another rank. Per-rank results of
        ORBmatcher.knnMatch2 on rows my_range(n_queries) (numpy or torch int32 [my_rows]) -> the four global arrays
        (numpy) on every rank. One all_gather of 16 B per query is the only communication."""
        def t(a):
            if isinstance(a, np.ndarray):
                a = torch.from_numpy(np.ascontiguousarray(a))
            return a.to(self.device)
        packed = torch.stack([t(idx1), t(d1), t(idx2), t(d2)], dim=1)  # [my_rows, 4]: one collective instead of four
        g = self._gather_padded(packed, n_queries).cpu().numpy()
        return g[:, 0].copy(), g[:, 1].copy(), g[:, 2].copy(), g[:, 3].copy()

    def gather_counts(self, counts):
        """counts: int32 tensor [k, my_units] on self.device -> [k, n_units_total_padded...] per rank stacked:
        returns a tensor [world, k, m] (m = largest shard). The 'trivial result gather' of bench.py."""
        if self.world == 1:
            return counts.unsqueeze(0)
        counts = counts.contiguous()
        out = torch.empty((self.world * counts.shape[0],) + tuple(counts.shape[1:]), dtype=counts.dtype,
                          device=counts.device)
        dist.all_gather_into_tensor(out, counts)  # concatenation along dim 0 (the form every backend accepts)
        return out.view((self.world,) + tuple(counts.shape))
