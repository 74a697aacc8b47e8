"""On-device data pipeline.  The reference feeds the step from ``DataLoader(dataset, shuffle=True,
num_workers=4, pin_memory=True)`` and one H2D copy per tensor and batch (core/train/train_pigan.py:114-121,
351-357) — about 1e5 samples/s, three orders of magnitude below the step.  Here the dataset is resident in HBM and a
batch is a row gather (``pigan_gather_rows``) by a device permutation; synthetic datasets are generated on the device
(``pigan_generate_spectra``, the formula of data_loader.py:62-80).

``DeviceLoader`` yields the reference's 5-tuple ``(spectrum, params_denorm, params_norm, metrics_denorm,
metrics_norm)`` as device tensors, so ``train_pigan(dataloader=DeviceLoader(...), ...)`` and
``pretrain_forward_model`` run unchanged.  Under data parallelism every rank draws the same permutation (same seed
and epoch) and takes its contiguous slice of every global batch.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from . import native
from .native import check, lib


def _dev(x: torch.Tensor, device) -> torch.Tensor:
    return x.to(device=device, dtype=torch.float32).contiguous()


def gather_rows(src: torch.Tensor, index: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = src[index[i]] for a contiguous 2-D (or 1-D) CUDA tensor of any dtype; index int64 on the device."""
    if not src.is_cuda or not src.is_contiguous():
        raise RuntimeError("gather_rows needs a contiguous CUDA tensor — the B200 path has no CPU fallback")
    if index.dtype != torch.int64 or not index.is_cuda or not index.is_contiguous():
        raise RuntimeError("gather_rows: index must be a contiguous int64 CUDA tensor")
    n = src.shape[0]
    row_bytes = src[0].numel() * src.element_size() if n else 0
    if out is None:
        out = torch.empty((index.numel(),) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    if index.numel() and n:
        check(lib.pigan_gather_rows(src.data_ptr(), n, row_bytes, index.data_ptr(), index.numel(), out.data_ptr(),
                                    None, native.current_stream()))
    return out


def generate_spectra(n: int, device, seed: int = 0, first_index: int = 0, noise_level: float = 0.1,
                     params_denorm: Optional[torch.Tensor] = None, frequency: Optional[torch.Tensor] = None,
                     apply_offset: bool = True, noise_dump: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None, params_out: Optional[torch.Tensor] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(spectrum [n,S], params_denorm [n,4]) from the reference's generator formula (data_loader.py:62-80).
    ``out`` / ``params_out``: preallocated fp32 CUDA buffers to fill (streaming generation without allocations)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("generate_spectra needs a CUDA device — the B200 path has no CPU fallback")
    if frequency is None:
        frequency = torch.linspace(0.5, 3.0, 250, dtype=torch.float32, device=device)
    frequency = _dev(frequency, device)
    S = frequency.numel()
    if out is None:
        out = torch.empty(n, S, device=device, dtype=torch.float32)
    elif out.shape != (n, S) or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous fp32 CUDA tensor [{n}, {S}]")
    if params_denorm is None:
        p_in = None
        p_out = params_out if params_out is not None else torch.empty(n, 4, device=device, dtype=torch.float32)
        if p_out.shape != (n, 4) or p_out.dtype != torch.float32 or not p_out.is_cuda or not p_out.is_contiguous():
            raise ValueError(f"params_out must be a contiguous fp32 CUDA tensor [{n}, 4]")
    else:
        p_in = p_out = _dev(params_denorm, device)
    check(lib.pigan_generate_spectra(native.ptr(p_in), None if p_in is not None else p_out.data_ptr(),
                                     frequency.data_ptr(), n, S, float(noise_level), int(seed), int(first_index),
                                     int(apply_offset), out.data_ptr(), native.ptr(noise_dump),
                                     native.current_stream()))
    return out, p_out


def rank_slice(count: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of a global batch of ``count`` rows that ``rank`` takes: near-equal contiguous slices, the last
    ranks may get fewer (or no) rows of a ragged final batch."""
    per = (count + world - 1) // world
    return min(rank * per, count), min((rank + 1) * per, count)


class DeviceDataset:
    """The tensors of a MetamaterialDataset (data_loader.py:124-147, 185-219) resident on the GPU."""

    def __init__(self, spectra, params_denorm, params_norm, metrics_denorm, metrics_norm, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceDataset needs a CUDA device — the B200 path has no CPU fallback")
        self.spectra = _dev(spectra, self.device)
        self.params_denorm = _dev(params_denorm, self.device)
        self.params_norm = _dev(params_norm, self.device)
        self.metrics_denorm = _dev(metrics_denorm, self.device)
        self.metrics_norm = _dev(metrics_norm, self.device)
        n = self.spectra.shape[0]
        for t in (self.params_denorm, self.params_norm, self.metrics_denorm, self.metrics_norm):
            if t.shape[0] != n:
                raise ValueError("all dataset tensors need the same number of rows")

    def __len__(self) -> int:
        return self.spectra.shape[0]

    @classmethod
    def from_dataset(cls, ds, device) -> "DeviceDataset":
        """From the reference-compatible MetamaterialDataset (attributes of data_loader.py:139-147)."""
        t = torch.as_tensor
        return cls(t(ds.spectra), t(ds.parameters), t(ds.normalized_parameters), t(ds.metrics),
                   t(ds.normalized_metrics), device)

    @classmethod
    def synthetic(cls, n: int, device, seed: int = 42, noise_level: float = 0.1, metrics: str = "physics"
                  ) -> "DeviceDataset":
        """n synthetic rows generated on the device (spectra from ``pigan_generate_spectra``).
        ``metrics="physics"``: the eight metric columns come from the spectra themselves (``physics.two_peak_metrics``,
        i.e. the physics kernel at the two band minima) and are normalised as the dataset does it — min/max over the
        non-NaN entries of each column, NaN -> 0.5 (data_loader.py:198-219); ``metric_ranges`` holds the ranges.
        ``metrics="uniform"``: placeholders in (0,1) (what bench.py's shape-only batches use)."""
        from . import physics
        spec, p = generate_spectra(n, device, seed=seed, noise_level=noise_level)
        ranges = {}
        if metrics == "uniform":
            g = torch.Generator(device=device)
            g.manual_seed(seed)
            md = mn = torch.rand(n, 8, generator=g, device=device, dtype=torch.float32)
        elif metrics == "physics":
            md = physics.two_peak_metrics(spec)["metrics"]
            mn = md.clone()
            for i, name in enumerate(physics.METRIC_NAMES):
                col = md[:, i]
                ok = col[~torch.isnan(col)]
                lo, hi = (float(ok.min()), float(ok.max())) if ok.numel() > 0 else (0.0, 1.0)
                ranges[name] = (lo, hi)
                mn[:, i] = (col - lo) / (hi - lo) if hi - lo > 1e-6 else 0.5
            mn[torch.isnan(mn)] = 0.5
        else:
            raise ValueError("metrics must be 'physics' or 'uniform'")
        ds = cls(spec, p, (p - 2.2) / 0.6 * 2.0 - 1.0, md, mn, device)
        ds.metric_ranges = ranges
        ds.metric_name_to_idx = {name: i for i, name in enumerate(physics.METRIC_NAMES)}
        ds.param_ranges = {k: (2.2, 2.8) for k in ("r1", "r2", "w", "g")}
        ds.frequencies = torch.linspace(0.5, 3.0, spec.shape[1], dtype=torch.float64).numpy()
        # with these attributes the object can stand in for the ``dataset`` argument of train_pigan
        # (train_pigan.py:132,162,165-166 read param_ranges, frequencies, metric_name_to_idx)
        return ds


class DeviceLoader:
    """Iterable over shuffled batches of a DeviceDataset; ``len()`` and ``batch_size`` as a DataLoader has them."""

    def __init__(self, dataset: DeviceDataset, batch_size: int, shuffle: bool = True, drop_last: bool = False,
                 seed: int = 0, rank: int = 0, world: int = 1):
        if batch_size < 1 or not (0 <= rank < world):
            raise ValueError("batch_size >= 1 and 0 <= rank < world required")
        self.ds, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.seed, self.rank, self.world, self.epoch = int(seed), int(rank), int(world), 0

    def __len__(self) -> int:
        gb = self.batch_size * self.world
        n = len(self.ds)
        return n // gb if self.drop_last else (n + gb - 1) // gb

    def batch_indices(self, epoch: int) -> Iterator[torch.Tensor]:
        """This rank's index vector of every global batch of ``epoch`` (device int64)."""
        n, dev = len(self.ds), self.ds.device
        if self.shuffle:
            g = torch.Generator(device=dev)
            g.manual_seed(self.seed * 1_000_003 + epoch)
            perm = torch.randperm(n, generator=g, device=dev)
        else:
            perm = torch.arange(n, device=dev)
        gb = self.batch_size * self.world
        for b in range(len(self)):
            chunk = perm[b * gb:(b + 1) * gb]
            if self.world > 1:
                # data parallel: every rank must see the SAME number of rows (the trainers take global batch =
                # local rows x world for BatchNorm statistics, loss means and the gradient scale, and a rank that
                # skipped a step would leave its peers waiting in the exchange kernels): a ragged last global batch is
                # cut to a multiple of the world size on all ranks alike, and dropped when that leaves < 2 rows each
                per = chunk.numel() // self.world
                if per < 2:
                    continue
                yield chunk[self.rank * per:(self.rank + 1) * per].contiguous()
                continue
            yield chunk.contiguous()

    def __iter__(self):
        ds = self.ds
        for idx in self.batch_indices(self.epoch):
            yield tuple(gather_rows(t, idx) for t in (ds.spectra, ds.params_denorm, ds.params_norm, ds.metrics_denorm,
                                                      ds.metrics_norm))
        self.epoch += 1
