"""madmom.processors protocol (Processor / SequentialProcessor / ParallelProcessor).

Same call protocol as madmom 0.16.1 ``madmom/processors.py`` so the front-end processors compose
exactly like the chains built in madmom.features.beats.RNNBeatProcessor (reached from
/root/reference/backend/app/services/grid/beats.py:74): ``proc(data, **kwargs)`` -> ``process``;
plain callables such as ``np.hstack`` are called without keyword arguments.
"""
from __future__ import annotations

from collections.abc import MutableSequence


class Processor(object):
    def process(self, data, **kwargs):
        raise NotImplementedError("Must be implemented by subclass.")

    def __call__(self, *args, **kwargs):
        return self.process(*args, **kwargs)


def _process(process_tuple):
    processor, data, kwargs = process_tuple
    if isinstance(processor, Processor) or getattr(processor, "_accepts_kwargs", False):
        return processor(data, **kwargs)
    return processor(data)


class SequentialProcessor(MutableSequence, Processor):
    def __init__(self, processors):
        self.processors = []
        for p in processors:
            if type(p) is SequentialProcessor:       # nested sequences are flattened
                self.processors.extend(p.processors)
            else:
                self.processors.append(p)

    def __getitem__(self, index):
        return self.processors[index]

    def __setitem__(self, index, processor):
        self.processors[index] = processor

    def __delitem__(self, index):
        del self.processors[index]

    def __len__(self):
        return len(self.processors)

    def insert(self, index, processor):
        self.processors.insert(index, processor)

    def process(self, data, **kwargs):
        for p in self.processors:
            data = _process((p, data, kwargs))
        return data


class ParallelProcessor(SequentialProcessor):
    """Branches run one after another on the host; their kernels overlap on the device."""

    def __init__(self, processors, num_threads=None):
        self.processors = list(processors)
        self.num_threads = 1 if num_threads is None else num_threads

    def process(self, data, **kwargs):
        return [_process((p, data, kwargs)) for p in self.processors]
