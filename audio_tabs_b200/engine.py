"""Chain fuser: turns a lazy madmom-shaped stage into kernel launches.

``run_chain(stage)`` walks ``stage -> ... -> ShortTimeFourierTransform -> FramedSignal -> Signal``,
builds the matching ``ResolutionSpec`` and calls the C ABI.  An intact chain is one fused launch
(b200spec_logfilt / b200spec_stft / b200spec_spectrogram); a chain rooted at a caller-supplied
host matrix uses the stand-alone K2/K3 kernels.  Nothing is computed on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi
from .plan import FrontEnd, Packed, ResolutionSpec, get_plan, require_cuda, _ptr, _stream_ptr


def _device_index() -> int:
    require_cuda(0)
    return torch.cuda.current_device()


def _signal_tensor(signal, device):
    """(tensor on device, dtype name). float32 / int16 pass through; float64 is narrowed to float32."""
    from .audio.signal import DeviceSignal
    if isinstance(signal, DeviceSignal):
        t = signal.data
        if t.device != device:
            t = t.to(device)
    else:
        arr = np.asarray(signal)
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)
        if arr.dtype not in (np.float32, np.int16):
            raise ValueError("signal dtype %s is not supported on the device (float32 or int16)" % arr.dtype)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=True)
    if t.dtype == torch.float64:
        t = t.to(torch.float32)
    if t.dtype == torch.float32:
        return t.contiguous(), "f32"
    if t.dtype == torch.int16:
        return t.contiguous(), "i16"
    raise ValueError("signal dtype %s is not supported on the device (float32 or int16)" % t.dtype)


def _parse(stage):
    """Collect the recipe from the last stage back to the root."""
    from .audio.spectrogram import (FilteredSpectrogram, LogarithmicSpectrogram, Spectrogram,
                                    SpectrogramDifference, StackedDifference)
    from .audio.stft import ShortTimeFourierTransform
    from .audio.chroma import FoldedChroma
    rec = dict(stack=False, diff=None, log=None, log_scale=1.0, filterbank=None, magnitude=False, stft=None, host=None,
               fold=None)
    node = stage
    if isinstance(node, FoldedChroma):
        rec["fold"] = (node.classes, node.num_classes)
        node = node.source
    if isinstance(node, StackedDifference):
        rec["stack"] = True
        node = node.diff
    if isinstance(node, SpectrogramDifference):
        rec["diff"] = (node.diff_frames, bool(node.positive_diffs), int(node.diff_max_bins or 0))
        node = node.source
    if isinstance(node, LogarithmicSpectrogram):
        rec["log"] = (float(node.mul), float(node.add) + float(getattr(node, "log_shift", 0.0)))
        rec["log_scale"] = float(getattr(node, "log_scale", 1.0))
        node = node.source
    if isinstance(node, FilteredSpectrogram):
        rec["filterbank"] = node.filterbank
        node = node.source
    if isinstance(node, Spectrogram):
        rec["magnitude"] = True
        node = node.source
    if isinstance(node, ShortTimeFourierTransform):
        rec["stft"] = node
    elif isinstance(node, np.ndarray):
        rec["host"] = node
    else:
        raise TypeError("cannot fuse a chain containing %s" % type(node))
    return rec


def _spec_from(rec, stft=None, frame_size=None):
    diff_frames, positive, max_bins = rec["diff"] if rec["diff"] else (0, False, 0)
    mul, add = rec["log"] if rec["log"] else (1.0, 1.0)
    fold = dict(proj_classes=rec["fold"][0], num_classes=rec["fold"][1]) if rec.get("fold") else {}
    if stft is not None:
        win, origin = stft.kernel_window_and_origin()
        return ResolutionSpec(frame_size=stft.fft_size, hop_size=stft.frames.hop_size,
                              origin=origin, fft_window=win, circular_shift=stft.circular_shift,
                              include_nyquist=stft.include_nyquist,
                              filterbank=rec["filterbank"], log=rec["log"] is not None, mul=mul, add=add,
                              log_scale=rec["log_scale"], diff_frames=diff_frames, positive_diffs=positive,
                              diff_max_bins=max_bins, **fold)
    return ResolutionSpec(frame_size=frame_size, filterbank=rec["filterbank"], log=rec["log"] is not None,
                          mul=mul, add=add, log_scale=rec["log_scale"], diff_frames=diff_frames, positive_diffs=positive,
                          diff_max_bins=max_bins)


def _frame_off(total, device):
    return torch.tensor([0, total], dtype=torch.int64, device=device)


def _standalone_tail(plan, spec, x, rec, device):
    """x: (T, K) float32 magnitudes on the device -> filter / log / diff with the K2/K3 kernels."""
    lib = _ffi.lib()
    T = x.shape[0]
    stream = _stream_ptr(None, device)
    if rec["filterbank"] is not None or rec["log"] is not None:
        width = spec.num_bands if rec["filterbank"] is not None else x.shape[1]
        y = torch.empty((T, width), dtype=torch.float32, device=device)
        _ffi.check(lib.b200spec_filter_log(plan.handle, 0, _ptr(x), x.shape[1], T,
                                           int(rec["filterbank"] is not None), int(rec["log"] is not None),
                                           _ptr(y), width, stream))
        x = y
    if rec["diff"] is None:
        return x
    B = x.shape[1]
    out = torch.empty((T, 2 * B if rec["stack"] else B), dtype=torch.float32, device=device)
    od = _ffi.OutDesc()
    od.d_out, od.ld_out = out.data_ptr(), out.shape[1]
    od.col_spec, od.col_diff = (0, B) if rec["stack"] else (-1, 0)
    fo = _frame_off(T, device)
    _ffi.check(lib.b200spec_diff_flux_chroma(plan.handle, 0, _ptr(x), B, _ptr(fo), 1, T, C.byref(od), stream))
    return out


def buffered_difference(buf, diff_frames, positive, max_bins, stacked):
    """Online SpectrogramDifferenceProcessor: the lagged difference of a device-resident buffer (diff_frames + T, B)
    whose first rows may be +inf ("no history yet": those differences are 0, madmom sets inf differences to 0).
    Returns rows diff_frames.. of [spec | diff] (stacked) or of the diff alone."""
    device = buf.device
    rec = dict(filterbank=None, log=None, log_scale=1.0, diff=(int(diff_frames), bool(positive), int(max_bins)),
               stack=bool(stacked))
    spec = _spec_from(rec, frame_size=2048)
    plan = get_plan(device.index, "f32", 1, [spec])
    out = _standalone_tail(plan, spec, buf.contiguous(), rec, device)[diff_frames:]
    B = buf.shape[1]
    d = out[:, -B:]
    d.masked_fill_(torch.isinf(d), 0.0)
    return out


def run_chain(stage, kind=None):
    """Materialise `stage` on the GPU; returns a torch tensor (complex64 for an STFT, else float32)."""
    rec = _parse(stage)
    device = torch.device("cuda", _device_index())
    lib = _ffi.lib()

    if rec["stft"] is None:
        # chain rooted at a host magnitude matrix
        host = rec["host"]
        x = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)).to(device)
        if rec["filterbank"] is None and rec["log"] is None and rec["diff"] is None:
            return x
        spec = _spec_from(rec, frame_size=2 * host.shape[1] if rec["filterbank"] is not None else 2048)
        plan = get_plan(device.index, "f32", 1, [spec])
        return _standalone_tail(plan, spec, x, rec, device)

    stft = rec["stft"]
    frames = stft.frames
    sig, dtype = _signal_tensor(frames.signal, device)
    if sig.ndim != 1:
        raise ValueError("frames must be a 2D array or iterable, got %s with shape %s." % (type(frames), frames.shape))
    spec = _spec_from(rec, stft=stft)
    packed = Packed(sig, [sig.shape[0]], spec.hop_size, num_frames=[frames.num_frames])

    norm = bool(getattr(frames.signal, "norm", False)) and not isinstance(frames.signal, np.ndarray)
    if rec["filterbank"] is not None:
        fe = FrontEnd([spec], device=device.index, dtype=dtype, channels=1)
        B = spec.num_bands
        scale = fe.peak_scales(packed, eps=0.0) if norm else None     # device signal with norm=True
        if rec["fold"] is not None:
            proj = torch.empty((packed.total_frames, spec.num_classes), dtype=torch.float32, device=device)
            fe.run_packed(packed, out=False, proj=[proj], clip_scale=scale)
            return proj
        if rec["diff"] is None or rec["stack"]:
            return fe.run_packed(packed, clip_scale=scale)   # [spec] or [spec | diff]
        full = fe.run_packed(packed, clip_scale=scale)       # diff only: second half of the stacked rows
        return full[:, B:].contiguous()

    if norm:
        raise ValueError("norm=True on a device-resident signal is applied as a gain inside the fused "
                         "filterbank chains; normalise the tensor yourself for a bare STFT / spectrogram")
    fe = FrontEnd([spec], device=device.index, dtype=dtype, channels=1)
    if not rec["magnitude"]:
        return fe.stft_packed(packed, 0, complex_out=True)
    x = fe.stft_packed(packed, 0, complex_out=False)
    if rec["log"] is None and rec["diff"] is None:
        return x
    return _standalone_tail(fe.plan, spec, x, rec, device)
