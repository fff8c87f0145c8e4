"""Reader / checker for the reference-fixture interchange files (test infrastructure).

tests/golden/reference_inputs.txt    written by tests/golden/make_reference_inputs.py
tests/golden/reference_fixtures.txt  written by oracle/ref_fixtures (Rust, over the REAL gym-rs crate)
                                     -- absent until someone runs it on a box with cargo + SDL2.
Format: oracle/ref_fixtures/src/main.rs.  Floats are 16-hex-digit f64 bit patterns.
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
INPUTS = os.path.join(HERE, "reference_inputs.txt")
FIXTURES = os.path.join(HERE, "reference_fixtures.txt")


def f(hexs: str) -> float:
    return struct.unpack("<d", struct.pack("<Q", int(hexs, 16)))[0]


def hx(v: float) -> str:
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(v)))[0]


def parse(path):
    """-> list of records: dict(kind='step'|'seq'|'reset', env, ..., inputs, outputs)"""
    recs, cur = [], None
    for raw in open(path):
        if raw.startswith("#") or not raw.strip():
            continue
        if raw.startswith("  "):                      # a step line of the running `seq` record
            t = raw.split()
            cur["steps"].append(dict(state=[f(x) for x in t[:4]], reward=f(t[4]), done=int(t[5])))
            continue
        left, _, right = raw.partition("->")
        t, o = left.split(), right.split()
        if t[0] == "step" and t[1] == "cartpole":
            cur = dict(kind="step", env="cartpole", integrator=int(t[2]), action=int(t[3]), state=[f(x) for x in t[4:8]])
            if o:
                cur.update(next_state=[f(x) for x in o[:4]], reward=f(o[4]), done=int(o[5]), truncated=int(o[6]))
        elif t[0] == "step" and t[1] == "mountain_car":
            cur = dict(kind="step", env="mountain_car", action=int(t[2]), state=[f(x) for x in t[3:5]])
            if o:
                cur.update(next_state=[f(x) for x in o[:2]], reward=f(o[2]), done=int(o[3]), truncated=int(o[4]))
        elif t[0] == "seq":
            cur = dict(kind="seq", env="cartpole", action=int(t[2]), n=int(t[3]), state=[f(x) for x in t[4:8]], steps=[])
        elif t[0] == "reset":
            cur = dict(kind="reset", env=t[1], seed=int(t[2]))
            if o:
                cur["state"] = [f(x) for x in o]
        else:
            raise ValueError("unknown line: " + raw)
        recs.append(cur)
    return recs


def ulps(a: float, b: float) -> int:
    if a == b:
        return 0
    ia, ib = (struct.unpack("<q", struct.pack("<d", v))[0] for v in (a, b))
    return abs(ia - ib) if (a > 0) == (b > 0) else 1 << 62


def check_against_oracle(recs, oracle, max_ulps=2):
    """Every record's outputs against the C oracle.  sin / cos come from the platform libm on both
    sides (Rust's f64::sin calls it too), so state values may differ in the last bits: <= max_ulps per
    step (accumulating along a `seq`); reward / done / truncated are exact.  Returns the number of
    values compared."""
    import numpy as np
    compared = 0
    for r in recs:
        if r["kind"] == "step":
            if r["env"] == "cartpole":
                p = oracle.default_params(oracle.CARTPOLE)
                p.kinematics_integrator = r["integrator"]
                got = oracle.step_batch(oracle.CARTPOLE, np.array(r["state"]).reshape(4, 1), [r["action"]],
                                        sbt=np.array([-1]), params=p)
            else:
                got = oracle.step_batch(oracle.MOUNTAIN_CAR, np.array(r["state"]).reshape(2, 1), [r["action"]])
            for k, want in enumerate(r["next_state"]):
                assert ulps(float(got["state"][k, 0]), want) <= max_ulps, (r, k, float(got["state"][k, 0]))
                compared += 1
            assert float(got["reward"][0]) == r["reward"] and int(got["done"][0]) == r["done"], r
            assert r["truncated"] == 0                         # cartpole.rs:480, mountain_car.rs:432
            compared += 2
        elif r["kind"] == "seq":
            st, sbt = np.array(r["state"]).reshape(4, 1), np.array([-1], dtype=np.int64)
            assert len(r["steps"]) == r["n"]
            for i, v in enumerate(r["steps"]):
                got = oracle.step_batch(oracle.CARTPOLE, st, [r["action"]], sbt=sbt)
                st, sbt = got["state"], got["sbt"]
                for k in range(4):
                    assert ulps(float(st[k, 0]), v["state"][k]) <= 4 * max_ulps * (i + 1), (i, k)
                assert float(got["reward"][0]) == v["reward"] and int(got["done"][0]) == v["done"], (i, v)
                compared += 6
        elif r["kind"] == "reset":
            # rand_pcg's stream is not reproduced (Philox here): range evidence only
            s = r["state"]
            if r["env"] == "cartpole":
                assert len(s) == 4 and all(-0.05 <= v < 0.05 for v in s), r       # cartpole.rs:353-363
            else:
                assert -0.6 <= s[0] < -0.4 and s[1] == 0.0, r                    # mountain_car.rs:162-189
            compared += len(s)
    return compared


def emulate_with_oracle(in_path, out_path, oracle):
    """Writes a fixtures file in the Rust program's format using the ORACLE as the producer.  Used only
    to test the reader / checker plumbing; a file made this way pins nothing."""
    import numpy as np
    out = ["# EMULATED by the oracle (plumbing test only)"]
    for r, raw in zip(parse(in_path), [l for l in open(in_path) if l.strip() and not l.startswith("#")]):
        raw = raw.strip()
        if r["kind"] == "step" and r["env"] == "cartpole":
            p = oracle.default_params(oracle.CARTPOLE)
            p.kinematics_integrator = r["integrator"]
            g = oracle.step_batch(oracle.CARTPOLE, np.array(r["state"]).reshape(4, 1), [r["action"]], sbt=np.array([-1]), params=p)
            out.append(raw + " -> " + " ".join(hx(v) for v in g["state"][:, 0]) + f" {hx(g['reward'][0])} {int(g['done'][0])} 0")
        elif r["kind"] == "step":
            g = oracle.step_batch(oracle.MOUNTAIN_CAR, np.array(r["state"]).reshape(2, 1), [r["action"]])
            out.append(raw + " -> " + " ".join(hx(v) for v in g["state"][:, 0]) + f" {hx(g['reward'][0])} {int(g['done'][0])} 0")
        elif r["kind"] == "seq":
            out.append(raw + " ->")
            st, sbt = np.array(r["state"]).reshape(4, 1), np.array([-1], dtype=np.int64)
            for _ in range(r["n"]):
                g = oracle.step_batch(oracle.CARTPOLE, st, [r["action"]], sbt=sbt)
                st, sbt = g["state"], g["sbt"]
                out.append("  " + " ".join(hx(v) for v in st[:, 0]) + f" {hx(g['reward'][0])} {int(g['done'][0])}")
        else:
            kind = oracle.CARTPOLE if r["env"] == "cartpole" else oracle.MOUNTAIN_CAR
            s = oracle.reset_batch(kind, 1, seed=r["seed"] & (2 ** 63 - 1))
            out.append(raw + " -> " + " ".join(hx(v) for v in s[:, 0]))
    with open(out_path, "w") as fh:
        fh.write("\n".join(out) + "\n")
