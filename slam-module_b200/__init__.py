"""B200-native ORB front-end + Hamming matching hot path of the SLAM module (host-side mirror of
the reference interfaces above the C-ABI library `csrc/libslamgpu.so`)."""
from . import synth  # noqa: F401
