"""CPU logic-check binding (TEST INFRASTRUCTURE): the library's own sources compiled against the SIMT emulator
(`tests/emu/cuda_emu.h` -> `tests/emu/libcdra_emu.so`) bound to the product's `Engine` so that the `not gpu` tests can
exercise plans, arenas, channel maps, BatchNorm bookkeeping and the agent API without a device.  The product package has no
switch that reaches this: the tests subclass the two GPU-specific hooks of `cdra.engine.Engine`."""
import os

from cdra import _lib
from cdra.engine import Engine

EMU_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libcdra_emu.so')


class EmuEngine(Engine):
    def __init__(self, batch, height=90, width=120, dtype='f32', image_u8=True, device='cpu', share=None):
        super().__init__(batch, height, width, dtype=dtype, image_u8=image_u8, device=device, share=share)

    def _open_library(self):
        return _lib.bind(EMU_PATH)

    def _check_device(self):
        assert self.device.type == 'cpu'

    def _stream(self):
        return None
