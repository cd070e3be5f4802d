"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): 5 streams, 2 chunks through the
synchronous step, the pipelined int16 step with the device-side phrase splitter, the feature-input step and the state
gather / scatter kernels.  Usage: compute-sanitizer --tool memcheck python tools/gpu_sanitize.py [B]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    M = tb.model
    eng = tb.Engine(tb.weights.init_weights(0), max_slots=2 * B + 2, max_batch=B)
    s1, s2 = eng.alloc_slots(B), eng.alloc_slots(B)
    pcm = tb.synth.telephony_pcm(B, 4800, seed=3)
    for i in range(2):
        c = pcm[:, i * 2400:(i + 1) * 2400]
        lp, tk = eng.step(s1, c)
        r = eng.wait(eng.submit(s2, c.astype(np.int16), M.OUT_LOGPROBS | M.OUT_PHRASES, np.full(B, i == 1, dtype=np.uint8)))
        assert np.array_equal(r["logprobs"], lp) and np.isfinite(lp).all()
    st = eng.export_states(s1)
    eng.import_states(s2, st)
    assert np.array_equal(eng.export_states(s2), st)
    eng.step_features(s2, np.zeros((B, 64, 30), dtype=np.float16))
    eng.reset_slots(s2)
    print(f"sanitize target ok: B={B}, launches per step {eng._get_info().launches_per_step}, phrases {len(r['phrases'])}")
    eng.close()


if __name__ == "__main__":
    main()
