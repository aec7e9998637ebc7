"""Host mirror of `Challenger` (plonky2/src/iop/challenger.rs:15-150): an overwrite-mode duplex sponge of rate 8 over
the configuration's permutation.  Challenges are popped from the END of the output buffer (:84-94), observing an
element invalidates buffered outputs (:41-50).

The permutation runs on the GPU (p2g_permute); a `permute` callable can be injected instead -- the tests build their
expected transcripts that way with the CPU oracle's permutation."""
import numpy as np

from . import _lib
from .fields import Field, REPR_CANONICAL

RATE = 8


class Challenger:
    def __init__(self, field: Field, ctx=None, permute=None):
        self.field = field
        self.ctx = ctx
        self._permute = permute
        self.sponge_state = np.zeros(field.width, dtype=field.dtype)
        self.input_buffer = []
        self.output_buffer = []

    # -- the permutation ------------------------------------------------------------------------------------------
    def _apply_permutation(self):
        if self._permute is not None:
            self.sponge_state = np.asarray(self._permute(self.sponge_state.copy()), dtype=self.field.dtype)
            return
        ctx = self.ctx or _lib.default_context()
        state = np.ascontiguousarray(self.sponge_state)
        ctx.check(ctx.lib.p2g_permute(ctx.handle, self.field.hash_id, REPR_CANONICAL, state.ctypes.data, 1))
        self.sponge_state = state

    # -- observing (:41-82) ----------------------------------------------------------------------------------------
    def observe_element(self, element):
        self.output_buffer = []
        self.input_buffer.append(int(element) % self.field.order)
        if len(self.input_buffer) == RATE:
            self.duplexing()

    def observe_elements(self, elements):
        for e in np.asarray(elements).reshape(-1):
            self.observe_element(e)

    def observe_extension_element(self, element):
        self.observe_elements(element)  # as_base_slice

    def observe_extension_elements(self, elements):
        self.observe_elements(elements)

    def observe_hash(self, hash_elements):
        self.observe_elements(hash_elements)

    def observe_cap(self, cap):
        hashes = cap.hashes if hasattr(cap, "hashes") else cap
        for h in hashes:
            self.observe_hash(h)

    # -- challenges (:84-124) --------------------------------------------------------------------------------------
    def get_challenge(self) -> int:
        if self.input_buffer or not self.output_buffer:
            self.duplexing()
        return self.output_buffer.pop()

    def get_n_challenges(self, n: int):
        return [self.get_challenge() for _ in range(n)]

    def get_extension_challenge(self):
        return np.array(self.get_n_challenges(self.field.extension_degree), dtype=self.field.dtype)

    def duplexing(self):
        assert len(self.input_buffer) <= RATE
        for i, x in enumerate(self.input_buffer):
            self.sponge_state[i] = x
        self.input_buffer = []
        self._apply_permutation()
        self.output_buffer = [int(x) for x in self.sponge_state[:RATE]]
