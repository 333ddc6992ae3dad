"""Drop-in equivalents of madmom.audio.{signal,stft,spectrogram,filters,chroma} for the hot path."""
from . import signal, stft, spectrogram, chroma  # noqa: F401
from .. import filters  # noqa: F401
