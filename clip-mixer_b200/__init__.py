"""B200-native Mixer-CLIP training hot path (drop-in for corentin-ryr/CLIP-mixer's
``training/clip`` model surface).  PyTorch is used for device memory, streams and
torch.distributed; all arithmetic on the path runs in hand-written sm_100a kernels behind the C
ABI of ``include/mixerclip.h`` (``libmixerclip.so``).  No CPU fallback."""

__version__ = "0.1.0"
