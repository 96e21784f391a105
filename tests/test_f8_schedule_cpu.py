"""Host logic of the f8c posterior kernel: the step schedule every role of k_posterior_fast8 walks
(optimobo_b200/csrc/posterior_fast8.cu::f8_build_schedule).  Pure host code -- no GPU needed: the library only has to load.

Invariants (for every n_pad the library accepts and several generator-time settings):
* chunk c (columns 256c .. of V = K* L^-T, reference: the triangular solve behind GPy's predict variance,
  util_functions.py:265 `model.predict`) is multiplied with exactly the K-blocks 0 .. min(4c+3, nkb-1), each once;
* its first step carries FIRST (accumulator reset), its last step DONE (epilogue drains), nothing in between;
* a TMEM slot holds one chunk at a time;
* fresh blocks come in increasing order, each exactly once; a reload never precedes the block's generation and
  the generated block carries STORE iff some later step reloads it."""
import ctypes

import numpy as np
import pytest

from optimobo_b200 import _cabi


def _schedule(n_pad, tg, max_run=0):
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    lib.ombo_debug_f8_schedule.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros(1 << 16, np.uint32)
    n = lib.ombo_debug_f8_schedule(n_pad, tg, max_run, buf.ctypes.data, buf.size)
    assert 0 < n <= buf.size
    return [int(w) for w in buf[:n]]


@pytest.mark.parametrize("max_run", [0, 1, 3])      # 0 = the pass schedule (default), >= 1 = rolling
@pytest.mark.parametrize("tg", [0.5, 3.0, 6.0])
@pytest.mark.parametrize("n_pad", [128, 256, 384, 512, 640, 768, 1024, 1152, 1536, 2048, 4096, 8192])
def test_schedule_invariants(n_pad, tg, max_run):
    steps = _schedule(n_pad, tg, max_run)
    nkb, n_chunks = n_pad // 64, (n_pad + 255) // 256
    seen = {c: [] for c in range(n_chunks)}
    state = {}                      # chunk -> "open" / "done"
    slot_chunk = [None, None]
    generated, reloaded, stored = [], set(), set()
    for w in steps:
        kb, fresh, store = w & 0xFF, (w >> 8) & 1, (w >> 9) & 1
        if fresh:
            assert kb == len(generated), "fresh blocks in order, once"
            generated.append(kb)
            if store:
                stored.add(kb)
        else:
            assert kb < len(generated), "reload before generation"
            assert kb in stored, "reload of a block that was not copied to the cache"
            assert not store
            reloaded.add(kb)
        for s in (0, 1):
            c = ((w >> (10 + 6 * s)) & 63) - 1
            first, done = (w >> (22 + s)) & 1, (w >> (24 + s)) & 1
            if c < 0:
                assert not first and not done
                continue
            if first:
                assert c not in state and slot_chunk[s] is None, "slot reused before its chunk completed"
                state[c] = "open"
                slot_chunk[s] = c
            assert state.get(c) == "open" and slot_chunk[s] == c
            assert kb not in seen[c]
            seen[c].append(kb)
            if done:
                state[c] = "done"
                slot_chunk[s] = None
    assert generated == list(range(nkb))
    assert stored == reloaded, "STORE exactly on the blocks that are reloaded"
    for c in range(n_chunks):
        assert state.get(c) == "done"
        assert sorted(seen[c]) == list(range(min(4 * c + 3, nkb - 1) + 1))


def test_pass_schedule_c5():
    """n_pad = 1024 (BASELINE C5), default schedule: blocks 0-7 are generated for chunks 0 and 1 and copied to the
    cache, blocks 8-15 alternate with the 8 reloads, and every reload feeds both open chunks (two MMA units)."""
    steps = _schedule(1024, 3.0, 0)
    assert len(steps) == 24
    reloads = [w for w in steps if not (w >> 8) & 1]
    assert len(reloads) == 8 and all(((w >> 10) & 63) == 3 and ((w >> 16) & 63) == 4 for w in reloads)
    fresh = [(w >> 8) & 1 for w in steps[8:]]
    assert fresh == [1, 0] * 8


def test_rolling_schedule_c5_is_generator_paced():
    """rolling schedule (experimental, OMBO_F8_MAXRUN=1): all 12 reloads fit into the MMA time the fresh steps leave
    over, none is left for after the last generated block, no step multiplies nothing, never two reloads in a row."""
    steps = _schedule(1024, 3.0, 1)
    assert sum(1 for w in steps if not (w >> 8) & 1) == 12
    assert (steps[-1] >> 8) & 1 and (steps[-1] & 0xFF) == 15
    assert all(((w >> 10) & 0xFFF) != 0 for w in steps)
    fresh = [(w >> 8) & 1 for w in steps]
    assert all(fresh[i] or fresh[i + 1] for i in range(len(fresh) - 1)), "never two reloads in a row"
