"""B200-native ORB front-end + Hamming matching hot path of the SLAM module (host-side mirror of
the reference interfaces above the C-ABI library `csrc/libslamgpu.so`)."""
from . import synth  # noqa: F401
# `slamgpu` (the ctypes binding of csrc/libslamgpu.so) is imported explicitly by its users so that
# importing the package never needs the built library.
