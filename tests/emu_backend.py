"""TEST INFRASTRUCTURE ONLY: binds tests/emu/libhelmnet_emu.so (the kernel sources compiled against the
fiber emulator in tests/emu/cuda_emu.h) behind the same ctypes wrapper the product uses, accepting CPU
tensors as "device" memory.  Lets the not-gpu suite drive the real host logic and kernel index math.
"""
import os
import subprocess

from helmnet_b200._lib import HelmnetLib

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
EMU_LIB = os.path.join(EMU_DIR, "libhelmnet_emu.so")


def build_emu():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
    return EMU_LIB


class EmuLib(HelmnetLib):
    requires_cuda = False

    def __init__(self):
        super().__init__(build_emu())

    def check_tensor(self, t, name="tensor"):
        assert not t.is_cuda

    def stream_for(self, device):
        return 0
