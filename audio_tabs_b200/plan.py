"""Plans and the batch engine: host-side layer between the madmom-shaped processors and the C ABI.

A ``ResolutionSpec`` describes one madmom chain (FramedSignalProcessor -> STFT -> filterbank -> log
-> difference).  ``get_plan`` caches device plans the way madmom's processors cache ``fft_window``
and ``filterbank``.  ``FrontEnd`` runs batches of clips through the fused kernels, either on
device-resident packed input (``run_packed``) or from host arrays with pinned, stream-overlapped
copies (``process_batch``).

PyTorch is used for device memory, streams and events only.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import threading
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _ffi
from .filters import Filterbank

SUPPORTED_FRAME_SIZES = (1024, 2048, 4096, 8192)
ONE_LAUNCH_DEFAULT = False      # FrontEnd(one_launch=None): see DESIGN.md section 4 for the measurement behind the default


def _digest(*arrays) -> str:
    h = hashlib.sha1()
    for a in arrays:
        if a is None:
            h.update(b"-")
        else:
            a = np.ascontiguousarray(a)
            h.update(str(a.dtype).encode())
            h.update(str(a.shape).encode())
            h.update(a.tobytes())
    return h.hexdigest()


@dataclass
class ResolutionSpec:
    frame_size: int
    hop_size: float = 441.0
    origin: int = 0
    fft_window: Optional[np.ndarray] = None     # float64/32, length frame_size (madmom's fft_window)
    filterbank: Optional[Filterbank] = None     # None -> STFT / magnitude only
    log: bool = True
    mul: float = 1.0
    add: float = 1.0
    diff_frames: int = 0
    positive_diffs: bool = False
    diff_max_bins: int = 0                      # SuperFlux maximum filter width (0 / 1 = none)
    power: bool = False                         # filterbank on |X|^2 (librosa melspectrogram) instead of |X|
    log_scale: float = 1.0                      # out = log_scale * log10(max(mul*y + add, log_floor))
    log_floor: float = 0.0                      # <= 0: no clamp (madmom); librosa power_to_db: amin = 1e-10
    circular_shift: bool = False                # STFT of the half-swapped frame: bin k times (-1)^k (complex output only)
    include_nyquist: bool = False               # STFT / magnitude rows carry frame_size/2 + 1 bins (no filterbank then)
    proj_classes: Optional[np.ndarray] = None   # per band class index (or -1), e.g. chroma fold
    proj_matrix: Optional[np.ndarray] = None    # or a dense (B, C) projection
    num_classes: int = 0
    _key: str = field(default="", repr=False)

    def __post_init__(self):
        self.frame_size = int(self.frame_size)
        self.hop_size = float(self.hop_size)
        self.origin = int(self.origin)
        if self.fft_window is None:
            self.fft_window = np.hanning(self.frame_size)
        win = np.asarray(self.fft_window)
        if win.shape != (self.frame_size,):
            raise ValueError("window must have frame_size elements")
        self.window32 = np.ascontiguousarray(win, dtype=np.float32)
        if self.filterbank is not None and not isinstance(self.filterbank, Filterbank):
            raise TypeError("not a Filterbank type or instance: %s" % self.filterbank)
        if self.filterbank is not None and self.include_nyquist:
            raise ValueError("include_nyquist=True: the filterbank kernels take frame_size/2 bins; "
                             "filter a spectrogram without the Nyquist bin")
        if self.filterbank is not None and self.filterbank.shape[0] != self.frame_size // 2:
            raise ValueError("filterbank must have frame_size/2 bins")
        self.num_bands = 0 if self.filterbank is None else int(self.filterbank.shape[1])
        # projection in CSR-by-class form
        self.proj_off = self.proj_band = self.proj_weight = None
        if self.proj_classes is not None:
            cls = np.asarray(self.proj_classes, dtype=np.int64)
            C_ = int(self.num_classes or (cls.max() + 1))
            order = [np.nonzero(cls == c)[0] for c in range(C_)]
            self.num_classes = C_
            self.proj_off = np.concatenate(([0], np.cumsum([len(o) for o in order]))).astype(np.int32)
            self.proj_band = (np.concatenate(order) if order else np.zeros(0)).astype(np.int32)
            self.proj_weight = np.ones(len(self.proj_band), np.float32)
        elif self.proj_matrix is not None:
            P = np.asarray(self.proj_matrix, dtype=np.float32)
            if P.shape[0] != self.num_bands:
                raise ValueError("projection must have num_bands rows")
            self.num_classes = P.shape[1]
            bands = [np.nonzero(P[:, c])[0] for c in range(P.shape[1])]
            self.proj_off = np.concatenate(([0], np.cumsum([len(b) for b in bands]))).astype(np.int32)
            self.proj_band = (np.concatenate(bands) if bands else np.zeros(0)).astype(np.int32)
            self.proj_weight = np.concatenate([P[b, c] for c, b in enumerate(bands)]).astype(np.float32) \
                if bands else np.zeros(0, np.float32)
        self._key = "|".join(map(str, (
            self.frame_size, repr(self.hop_size), self.origin, _digest(self.window32),
            _digest(None if self.filterbank is None else np.asarray(self.filterbank)),
            int(self.log), repr(float(self.mul)), repr(float(self.add)), self.diff_frames,
            int(self.positive_diffs), int(self.diff_max_bins or 0), int(bool(self.power)),
            repr(float(self.log_scale)), repr(float(self.log_floor)), int(bool(self.circular_shift)), int(bool(self.include_nyquist)),
            _digest(self.proj_off, self.proj_band, self.proj_weight))))

    @property
    def num_bins(self):
        return self.frame_size // 2

    @property
    def out_width(self):
        """columns this resolution contributes to a stacked output: [spec | diff]"""
        return self.num_bands * (2 if self.diff_frames > 0 else 1)


class DevicePlan:
    """Owns one b200spec_plan (device constants for up to 4 resolutions)."""

    def __init__(self, device: int, dtype: str, channels: int, specs: Sequence[ResolutionSpec]):
        if len(specs) < 1 or len(specs) > _ffi.MAX_RES:
            raise ValueError("a plan holds 1..%d resolutions" % _ffi.MAX_RES)
        self.device, self.dtype, self.channels, self.specs = int(device), dtype, int(channels), list(specs)
        lib = _ffi.lib()
        desc = _ffi.PlanDesc()
        desc.device = self.device
        desc.dtype = {"f32": _ffi.F32, "i16": _ffi.I16}[dtype]
        desc.channels = self.channels
        desc.num_res = len(specs)
        keep = []

        def fptr(a):
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            return a.ctypes.data_as(_ffi.c_float_p)

        def iptr(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            keep.append(a)
            return a.ctypes.data_as(_ffi.c_int32_p)

        for i, s in enumerate(specs):
            r = desc.res[i]
            r.frame_size, r.hop_size, r.origin = s.frame_size, s.hop_size, s.origin
            r.window = fptr(s.window32)
            r.num_bands = s.num_bands
            if s.filterbank is not None:
                start, length, woff, weights = s.filterbank.banded()
                r.band_start, r.band_len, r.band_woff, r.weights = iptr(start), iptr(length), iptr(woff), fptr(weights)
            r.log_enabled = int(bool(s.log))
            r.mul, r.add = float(s.mul), float(s.add)
            r.diff_frames, r.positive_diffs = int(s.diff_frames), int(bool(s.positive_diffs))
            r.diff_max_bins = int(s.diff_max_bins or 0)
            r.power, r.log_scale, r.log_floor = int(bool(s.power)), float(s.log_scale), float(s.log_floor)
            r.circular_shift = int(bool(s.circular_shift))
            r.include_nyquist = int(bool(s.include_nyquist))
            r.num_classes = int(s.num_classes) if s.proj_off is not None else 0
            if s.proj_off is not None:
                r.proj_off, r.proj_band, r.proj_weight = iptr(s.proj_off), iptr(s.proj_band), fptr(s.proj_weight)
        handle = C.c_void_p()
        _ffi.check(lib.b200spec_plan_create(C.byref(desc), C.byref(handle)))
        self._handle = handle
        self._lib = lib

    @property
    def handle(self):
        return self._handle

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.b200spec_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass


_PLAN_CACHE = {}
_PLAN_LOCK = threading.Lock()


def get_plan(device: int, dtype: str, channels: int, specs: Sequence[ResolutionSpec]) -> DevicePlan:
    key = (int(device), dtype, int(channels), tuple(s._key for s in specs))
    with _PLAN_LOCK:
        plan = _PLAN_CACHE.get(key)
        if plan is None:
            plan = DevicePlan(device, dtype, channels, specs)
            _PLAN_CACHE[key] = plan
        return plan


def clear_plan_cache():
    with _PLAN_LOCK:
        for p in _PLAN_CACHE.values():
            p.close()
        _PLAN_CACHE.clear()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream: Optional[torch.cuda.Stream], device) -> C.c_void_p:
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)


def require_cuda(device: int = 0) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("audio_tabs_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", device)


def dtype_name(dt) -> str:
    dt = np.dtype(dt) if not isinstance(dt, torch.dtype) else dt
    if dt in (np.dtype(np.float32), torch.float32):
        return "f32"
    if dt in (np.dtype(np.int16), torch.int16):
        return "i16"
    raise ValueError("only float32 and int16 signals are supported on the device (got %s)" % dt)


class Packed:
    """A batch of clips packed back to back on the device, with per-clip sample/row offsets."""

    def __init__(self, sig: torch.Tensor, lengths: Sequence[int], hop_size: float, end: str = "normal",
                 num_frames: Optional[Sequence[int]] = None):
        self.sig = sig
        self.lengths = [int(n) for n in lengths]
        self.n_clips = len(self.lengths)
        if num_frames is None:
            num_frames = [_ffi.num_frames(n, hop_size, end) for n in self.lengths]
        self.num_frames = [int(t) for t in num_frames]
        clip_off = np.zeros(self.n_clips + 1, np.int64)
        np.cumsum(self.lengths, out=clip_off[1:])
        frame_off = np.zeros(self.n_clips + 1, np.int64)
        np.cumsum(self.num_frames, out=frame_off[1:])
        self.clip_off_host, self.frame_off_host = clip_off, frame_off
        self.total_frames = int(frame_off[-1])
        dev = sig.device
        self.clip_off = torch.from_numpy(clip_off).to(dev)
        self.frame_off = torch.from_numpy(frame_off).to(dev)


class FrontEnd:
    """Fused multi-resolution front end for one device.

    Output layout per frame row: for every resolution in order ``[spec_r | diff_r]`` (the
    ``np.hstack`` order of madmom's RNNBeatProcessor pre-processor); optional flux / projection
    outputs are separate matrices.
    """

    def __init__(self, specs: Sequence[ResolutionSpec], device: int = 0, dtype: str = "f32", channels: int = 1,
                 end: str = "normal", concurrent_streams: bool = False, one_launch: Optional[bool] = None):
        self.specs = list(specs)
        hops = {s.hop_size for s in self.specs}
        if len(hops) != 1:
            raise ValueError("all resolutions of a FrontEnd must share hop_size (rows are stacked per frame)")
        self.hop_size = self.specs[0].hop_size
        self.end = end
        self.device = require_cuda(device)
        self.dtype, self.channels = dtype, int(channels)
        self.plan = get_plan(device, dtype, channels, self.specs)
        self.col = []
        c = 0
        for s in self.specs:
            self.col.append(c)
            c += s.out_width
        self.width = c
        self._lib = _ffi.lib()
        self._workspaces = {}
        self._streams = None
        self._events = None
        self.concurrent_streams = concurrent_streams and len(self.specs) > 1
        # all resolutions in ONE launch (b200spec_logfilt_multi: each chunk's samples leave HBM once): needs 2-3
        # resolutions of frame size <= 4096 whose tables fit one SM together, and no SuperFlux second pass.
        # None = the library default (ONE_LAUNCH_DEFAULT); True raises if the plan cannot run that way.
        res_idx = (C.c_int32 * len(self.specs))(*range(len(self.specs)))
        can = (1 < len(self.specs) <= 3 and not any((s.diff_max_bins or 0) > 1 and s.diff_frames > 0 for s in self.specs)
               and bool(self._lib.b200spec_logfilt_multi_supported(self.plan.handle, len(self.specs), res_idx)))
        if one_launch and not can:
            raise ValueError("these resolutions cannot run in one launch (2-3 resolutions of frame size <= 4096 sharing "
                             "hop_size, tables within one SM's shared memory, no diff_max_bins)")
        self.one_launch = can and (ONE_LAUNCH_DEFAULT if one_launch is None else bool(one_launch))

    # ---- helpers ---------------------------------------------------------------------------
    def _workspace(self, res: int, n_clips: int) -> torch.Tensor:
        need = int(self._lib.b200spec_workspace_bytes(n_clips))
        ws = self._workspaces.get(res)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 4096), dtype=torch.uint8, device=self.device)
            self._workspaces[res] = ws
        return ws

    def _side_streams(self):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(self.device) for _ in self.specs[1:]]
            self._events = [torch.cuda.Event() for _ in self.specs]
        return self._streams

    def torch_dtype(self):
        return torch.float32 if self.dtype == "f32" else torch.int16

    def pack(self, signals: Sequence, non_blocking: bool = True) -> Packed:
        """Copy a list of host arrays / tensors into one packed device buffer."""
        tdt = self.torch_dtype()
        lens = [int(s.shape[0]) for s in signals]
        ch = self.channels
        total = sum(lens)
        shape = (total,) if ch == 1 else (total, ch)
        sig = torch.empty(shape, dtype=tdt, device=self.device)
        o = 0
        for s, n in zip(signals, lens):
            t = s if isinstance(s, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(s))
            if t.dtype != tdt:
                raise ValueError("signal dtype %s does not match the plan dtype %s" % (t.dtype, tdt))
            if (t.ndim == 2) != (ch == 2) or (t.ndim == 2 and t.shape[1] != ch):
                raise ValueError("signal shape %s does not match channels=%d" % (tuple(t.shape), ch))
            sig[o:o + n].copy_(t, non_blocking=non_blocking)
            o += n
        return Packed(sig, lens, self.hop_size, self.end)

    def alloc_output(self, total_frames: int) -> torch.Tensor:
        return torch.empty((total_frames, self.width), dtype=torch.float32, device=self.device)

    # ---- device-resident hot path ----------------------------------------------------------
    def run_packed(self, packed: Packed, out: Optional[torch.Tensor] = None, flux: Optional[List] = None,
                   proj: Optional[List] = None, timing: Optional[List] = None,
                   clip_scale: Optional[torch.Tensor] = None,
                   clip_status: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Launch the fused kernels for every resolution; returns the stacked (rows, width) tensor.

        ``out=False`` skips the stacked matrix (only ``flux`` / ``proj`` are written).
        ``clip_scale``: (n_clips,) float32 device tensor of per-clip gains applied before the logarithm
        (see :meth:`peak_scales`); ``timing``: a list that receives ``(resolution, start_event, end_event)`` per launch, recorded on
        the stream the kernel runs on (read them after a synchronise); ``clip_status``: (n_clips,) int32 device
        tensor, zeroed by the caller, that receives per-clip status bits (``_ffi.CLIP_NONFINITE``: NaN / Inf
        reached the clip's rows -- the other clips of the batch are unaffected).
        No synchronisation: results are ordered on the current stream.
        """
        if clip_status is not None and (clip_status.dtype != torch.int32 or clip_status.numel() < packed.n_clips
                                        or not clip_status.is_contiguous()):
            raise ValueError("clip_status must be a contiguous int32 tensor with n_clips elements")
        if out is None:
            out = self.alloc_output(packed.total_frames)
        if out is not False and (out.shape != (packed.total_frames, self.width) or out.dtype != torch.float32
                                 or not out.is_contiguous()):
            raise ValueError("out must be a contiguous float32 (total_frames, %d) tensor" % self.width)
        cur = torch.cuda.current_stream(self.device)
        if self.one_launch:
            n = len(self.specs)
            ods = (_ffi.OutDesc * n)()
            for r, s in enumerate(self.specs):
                od = ods[r]
                od.d_out = out.data_ptr() if out is not False else None
                od.ld_out = self.width
                od.col_spec = self.col[r]
                od.col_diff = self.col[r] + s.num_bands if s.diff_frames > 0 else -1
                od.d_flux = flux[r].data_ptr() if flux is not None and flux[r] is not None else None
                od.d_proj = proj[r].data_ptr() if proj is not None and proj[r] is not None else None
                od.ld_proj = proj[r].shape[1] if proj is not None and proj[r] is not None else 0
                od.d_clip_scale = clip_scale.data_ptr() if clip_scale is not None else None
                od.d_clip_status = clip_status.data_ptr() if clip_status is not None else None
            ws = self._workspace(0, packed.n_clips)
            if timing is not None:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(cur)
            _ffi.check(self._lib.b200spec_logfilt_multi(
                self.plan.handle, n, (C.c_int32 * n)(*range(n)), _ptr(packed.sig), _ptr(packed.clip_off),
                _ptr(packed.frame_off), packed.n_clips, packed.total_frames, ods, _ptr(ws), ws.numel(),
                C.c_void_p(cur.cuda_stream)))
            if timing is not None:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(cur)
                timing.append((-1, t0, t1))
            return out
        use_side = self.concurrent_streams and packed.total_frames > 0
        side = self._side_streams() if use_side else None
        if use_side:
            self._events[0].record(cur)
        for r, s in enumerate(self.specs):
            stream = cur
            if use_side and r > 0:
                stream = side[r - 1]
                stream.wait_event(self._events[0])
            od = _ffi.OutDesc()
            od.d_out = out.data_ptr() if out is not False else None
            od.ld_out = self.width
            od.col_spec = self.col[r]
            superflux = s.diff_frames > 0 and (s.diff_max_bins or 0) > 1
            want_flux = flux is not None and flux[r] is not None
            od.col_diff = self.col[r] + s.num_bands if s.diff_frames > 0 and not superflux else -1
            od.d_flux = flux[r].data_ptr() if want_flux and not superflux else None
            tmp_L = None
            if superflux and out is False:      # the SuperFlux pass needs the log-filtered rows somewhere
                tmp_L = torch.empty((packed.total_frames, s.num_bands), dtype=torch.float32, device=self.device)
                od.d_out, od.ld_out, od.col_spec = tmp_L.data_ptr(), s.num_bands, 0
            od.d_proj = proj[r].data_ptr() if proj is not None and proj[r] is not None else None
            od.ld_proj = proj[r].shape[1] if proj is not None and proj[r] is not None else 0
            od.d_clip_scale = clip_scale.data_ptr() if clip_scale is not None else None
            od.d_clip_status = clip_status.data_ptr() if clip_status is not None else None
            ws = self._workspace(r, packed.n_clips)
            if timing is not None:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(stream)
            _ffi.check(self._lib.b200spec_logfilt(
                self.plan.handle, r, _ptr(packed.sig), _ptr(packed.clip_off), _ptr(packed.frame_off),
                packed.n_clips, packed.total_frames, C.byref(od), _ptr(ws), ws.numel(),
                C.c_void_p(stream.cuda_stream)))
            if superflux:
                # SuperFlux: D[n] = L[n] - maxfilter(L[n-k]) needs whole lagged rows, so it runs as a
                # second kernel over the rows just written (they are still in L2)
                sd = _ffi.OutDesc()
                if tmp_L is not None:
                    src, ld_src = tmp_L, s.num_bands
                    sd.d_out, sd.ld_out, sd.col_spec, sd.col_diff = None, 0, -1, -1
                else:
                    src, ld_src = out[:, self.col[r]:], self.width
                    sd.d_out, sd.ld_out, sd.col_spec, sd.col_diff = out.data_ptr(), self.width, -1, self.col[r] + s.num_bands
                sd.d_flux = flux[r].data_ptr() if want_flux else None
                sd.d_proj, sd.ld_proj = None, 0
                _ffi.check(self._lib.b200spec_diff_flux_chroma(
                    self.plan.handle, r, C.c_void_p(src.data_ptr()), ld_src, _ptr(packed.frame_off), packed.n_clips,
                    packed.total_frames, C.byref(sd), C.c_void_p(stream.cuda_stream)))
            if timing is not None:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(stream)
                timing.append((r, t0, t1))
            if use_side and r > 0:
                self._events[r].record(stream)
                cur.wait_event(self._events[r])
        return out

    def peak_scales(self, packed: Packed, eps: float = 1e-9, reciprocal: bool = True) -> torch.Tensor:
        """Per-clip ``1 / (max|x| + eps)`` (or the peak itself) computed on the device in one read of the
        samples: the gain that makes ``run_packed(..., clip_scale=...)`` return the spectrogram of the
        peak-normalised clip (/root/reference/backend/app/services/audio.py:24-26; madmom ``norm=True``
        is ``eps=0``)."""
        out = torch.empty(packed.n_clips, dtype=torch.float32, device=self.device)
        _ffi.check(self._lib.b200spec_clip_peak(self.plan.handle, _ptr(packed.sig), _ptr(packed.clip_off),
                                                packed.n_clips, float(eps), int(bool(reciprocal)), _ptr(out),
                                                _stream_ptr(None, self.device)))
        return out

    def stft_packed(self, packed: Packed, res: int = 0, complex_out: bool = True) -> torch.Tensor:
        s = self.specs[res]
        if complex_out:
            out = torch.empty((packed.total_frames, s.num_bins + int(s.include_nyquist)), dtype=torch.complex64,
                              device=self.device)
            fn = self._lib.b200spec_stft
        else:
            out = torch.empty((packed.total_frames, s.num_bins + int(s.include_nyquist)), dtype=torch.float32,
                              device=self.device)
            fn = self._lib.b200spec_spectrogram
        ws = self._workspace(res, packed.n_clips)
        _ffi.check(fn(self.plan.handle, res, _ptr(packed.sig), _ptr(packed.clip_off), _ptr(packed.frame_off),
                      packed.n_clips, packed.total_frames, _ptr(out), _ptr(ws), ws.numel(),
                      _stream_ptr(None, self.device)))
        return out

    # ---- pinned host in / pinned host out, copies overlapped with compute ----------------------
    def _pipeline(self, lengths, group_clips, n_slots=2):
        """``group_clips``: clips per group, or a schedule of group sizes whose last entry repeats
        (``[1, 2]`` = a short first group, then pairs)."""
        sched = [int(group_clips)] if isinstance(group_clips, (int, np.integer)) else [int(v) for v in group_clips]
        if not sched or min(sched) < 1:
            raise ValueError("group sizes must be >= 1")
        key = (tuple(int(n) for n in lengths), tuple(sched), int(n_slots))
        cache = getattr(self, "_pipe_cache", None)
        if cache is not None and cache["key"] == key:
            return cache
        groups = []
        samp0 = row0 = 0
        bounds, g0 = [], 0
        while g0 < len(lengths):
            size = sched[min(len(bounds), len(sched) - 1)]
            bounds.append((g0, min(len(lengths), g0 + size)))
            g0 += size
        for g0, g1 in bounds:
            lens = [int(n) for n in lengths[g0:g1]]
            frames = [_ffi.num_frames(n, self.hop_size, self.end) for n in lens]
            groups.append(dict(lens=lens, frames=frames, samp0=samp0, nsamp=sum(lens), row0=row0, rows=sum(frames)))
            samp0 += sum(lens)
            row0 += sum(frames)
        max_s = max(g["nsamp"] for g in groups)
        max_r = max(g["rows"] for g in groups)
        sshape = (max_s,) if self.channels == 1 else (max_s, self.channels)
        slots = [dict(sig=torch.empty(sshape, dtype=self.torch_dtype(), device=self.device),
                      out=torch.empty((max_r, self.width), dtype=torch.float32, device=self.device))
                 for _ in range(n_slots)]
        for g in groups:  # per-group offsets live on the device for the whole pipeline lifetime
            g["packed"] = [Packed(slots[s]["sig"][:g["nsamp"]], g["lens"], self.hop_size, self.end, g["frames"])
                           for s in range(n_slots)]
        cache = dict(key=key, groups=groups, slots=slots, total_rows=row0, total_samples=samp0,
                     h2d=torch.cuda.Stream(self.device), comp=torch.cuda.Stream(self.device),
                     d2h=torch.cuda.Stream(self.device))
        self._pipe_cache = cache
        return cache

    def process_batch_pinned(self, host_in: torch.Tensor, lengths: Sequence[int], host_out: torch.Tensor,
                             group_clips=(1, 2), n_slots: int = 2) -> torch.Tensor:
        """Whole batch from a pinned packed host tensor to a pinned host result matrix.

        Clips are processed in groups; the host->device copy of group g+1 and the device->host copy
        of group g-1 overlap the kernels of group g on three streams with two device slots.  The
        current stream is joined at the end (but not synchronised with the host).  Small groups keep
        the fill and drain of the pipeline short, but every group costs a set of launches; the default
        schedule is one clip first (the only copy nothing overlaps), then pairs -- with an even number of
        clips the last group is a single clip again, which shortens the drain.  Measured on B200, 64 x 3-min
        stems, one box: (1, 2) 280.5 k audio-s/s, 2 clips per group 271.8 k, 1: 266 k, 4: 269.8 k, 8: 258 k;
        a third slot changes nothing (tools/e2e_sweep.py).
        """
        pipe = self._pipeline(lengths, group_clips, n_slots)
        if host_in.shape[0] != pipe["total_samples"] or tuple(host_out.shape) != (pipe["total_rows"], self.width):
            raise ValueError("host_in / host_out shapes do not match `lengths`")
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        h2d, comp, d2h = pipe["h2d"], pipe["comp"], pipe["d2h"]
        for s in (h2d, comp, d2h):
            s.wait_event(start)
        ev_comp = [None] * n_slots   # last compute on each slot (input slot free for the next H2D)
        ev_d2h = [None] * n_slots    # last D2H on each slot (output slot free for the next compute)
        for gi, g in enumerate(pipe["groups"]):
            slot = gi % n_slots
            buf = pipe["slots"][slot]
            if ev_comp[slot] is not None:
                h2d.wait_event(ev_comp[slot])
            with torch.cuda.stream(h2d):
                buf["sig"][:g["nsamp"]].copy_(host_in[g["samp0"]:g["samp0"] + g["nsamp"]], non_blocking=True)
                ev_h2d = torch.cuda.Event()
                ev_h2d.record(h2d)
            comp.wait_event(ev_h2d)
            if ev_d2h[slot] is not None:
                comp.wait_event(ev_d2h[slot])
            with torch.cuda.stream(comp):
                self.run_packed(g["packed"][slot], buf["out"][:g["rows"]])
                ev_comp[slot] = torch.cuda.Event()
                ev_comp[slot].record(comp)
            d2h.wait_event(ev_comp[slot])
            with torch.cuda.stream(d2h):
                host_out[g["row0"]:g["row0"] + g["rows"]].copy_(buf["out"][:g["rows"]], non_blocking=True)
                ev_d2h[slot] = torch.cuda.Event()
                ev_d2h[slot].record(d2h)
        for ev in ev_d2h:
            if ev is not None:
                cur.wait_event(ev)
        return host_out

    # ---- host in / host out ----------------------------------------------------------------
    def process_batch(self, signals: Sequence, return_tensors: bool = False, peak_normalize: bool = False,
                      eps: float = 1e-9, return_status: bool = False):
        """signals: list of host arrays (float32 / int16; (N,) or (N, 2)). Returns one (T_i, width) per clip.

        ``peak_normalize``: each clip is treated as ``y / (max|y| + eps)`` -- what the reference does before
        this path (/root/reference/backend/app/services/audio.py:24-26) -- fused as a per-clip gain.
        ``return_status``: also return a (n_clips,) int32 numpy array of per-clip status bits (0 = fine,
        ``_ffi.CLIP_NONFINITE`` = the clip's rows contain NaN / Inf because its samples did): one bad clip does
        not poison the batch -- the reference's analogue is one failed Celery task
        (/root/reference/backend/app/workers/tasks.py:35-38)."""
        packed = self.pack(signals)
        scale = self.peak_scales(packed, eps=eps) if peak_normalize else None
        status = torch.zeros(max(packed.n_clips, 1), dtype=torch.int32, device=self.device) if return_status else None
        out = self.run_packed(packed, clip_scale=scale, clip_status=status)
        if return_tensors:
            res = [out[packed.frame_off_host[i]:packed.frame_off_host[i + 1]] for i in range(packed.n_clips)]
            return (res, status[:packed.n_clips]) if return_status else res
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        st = status[:packed.n_clips].cpu().numpy() if return_status else None     # synchronises the stream
        torch.cuda.current_stream(self.device).synchronize()
        arr = host.numpy()
        res = [arr[packed.frame_off_host[i]:packed.frame_off_host[i + 1]] for i in range(packed.n_clips)]
        return (res, st) if return_status else res
