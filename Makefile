# Convenience targets; everything is also reachable through __graft_entry__.py / pytest / bench.py.
PY ?= python

build:            ## nvcc (sm_100a) + gcc oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle vs golden vectors, boundary, host logic (no GPU needed)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity tests through the C ABI on a B200
	$(PY) -m pytest tests -x -q -m gpu

smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## one JSON line
	$(PY) bench.py

bench-reference:  ## the reference-equivalent scalar CPU loop
	$(PY) bench.py --impl reference

golden:           ## regenerate tests/golden/step_vectors.json (mpmath)
	$(PY) tests/golden/make_golden.py

clean:
	rm -rf gym_rs_b200/libgymrs_b200.so gym_rs_b200/csrc/build oracle/libgymrs_oracle.so tests/cpp/env_test

.PHONY: build test test-gpu smoke bench bench-reference golden clean
