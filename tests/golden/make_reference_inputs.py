#!/usr/bin/env python
"""Write tests/golden/reference_inputs.txt -- the inputs oracle/ref_fixtures (a Rust program over the
REAL gym-rs crate) turns into tests/golden/reference_fixtures.txt.

Inputs: every (state, action) of tests/golden/step_vectors.json (SURVEY.md Appendix B cases included),
512 more seeded random pairs per env over the parity ranges of SURVEY.md section 8d, the 60-step
no-reset CartPole sequence (cartpole.rs:455-464), and a few seeded resets.  Floats are written as the
16-hex-digit bit pattern of the f64 (exact).  Format: oracle/ref_fixtures/src/main.rs.

Run:  python tests/golden/make_reference_inputs.py        (CPU only, deterministic)
"""
import json
import os
import random
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(v: float) -> str:
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def lines():
    g = json.load(open(os.path.join(HERE, "step_vectors.json")))
    out = ["# inputs for oracle/ref_fixtures (see tests/golden/make_reference_inputs.py)"]
    rnd = random.Random(20261018)
    cp = [(v["state"], v["action"]) for v in g["cartpole"]]
    cp += [([rnd.uniform(-2.4, 2.4), rnd.uniform(-3, 3), rnd.uniform(-0.21, 0.21), rnd.uniform(-3, 3)],
            rnd.randrange(2)) for _ in range(512)]
    for integ in (0, 1):
        for s, a in cp:
            out.append(f"step cartpole {integ} {a} " + " ".join(hx(x) for x in s))
    mc = [(v["state"], v["action"]) for v in g["mountain_car"]]
    mc += [([rnd.uniform(-1.2, 0.6), rnd.uniform(-0.07, 0.07)], rnd.randrange(3)) for _ in range(512)]
    for s, a in mc:
        out.append(f"step mountain_car {a} " + " ".join(hx(x) for x in s))
    out.append("seq cartpole 1 60 " + " ".join(hx(0.0) for _ in range(4)))
    out.append("seq cartpole 0 80 " + " ".join(hx(x) for x in (0.5, -0.25, 0.05, 0.1)))
    for seed in (0, 1, 42, 64, 2 ** 63 + 12345):
        out.append(f"reset cartpole {seed}")
        out.append(f"reset mountain_car {seed}")
    return out


if __name__ == "__main__":
    p = os.path.join(HERE, "reference_inputs.txt")
    with open(p, "w") as f:
        f.write("\n".join(lines()) + "\n")
    print("wrote", p)
