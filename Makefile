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

gpu-check:        ## on a GPU box: parity tests, smoke, one bench line per env (set GYMRS_CHECK_ENVS)
	bash tools/gpu_check.sh

profile:          ## on a GPU box: the ncu passes behind profiles/ (then: python profiles/summarize.py r01)
	bash tools/profile_round.sh

sanitize:         ## on a GPU box: compute-sanitizer memcheck / racecheck / initcheck / synccheck
	bash tools/sanitize_round.sh

mode-floor: build ## tools/bin/mode_floor: copy-only floors of every stepping mode (run it on a GPU box)
	mkdir -p tools/bin
	nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/bin/mode_floor tools/mode_floor.cu \
	    -Iinclude -Lgym_rs_b200 -lgymrs_b200 -Xlinker -rpath='$$ORIGIN/../../gym_rs_b200'

scale-check:      ## on an 8-GPU box: the driver's 1/2/4/8-GPU sequence, both arms
	bash tools/scale_check.sh

clean:
	rm -rf gym_rs_b200/libgymrs_b200.so gym_rs_b200/csrc/build oracle/libgymrs_oracle.so tests/cpp/env_test \
	       tests/cuda/curand_check tools/bin

.PHONY: build test test-gpu smoke bench bench-reference golden gpu-check profile sanitize mode-floor scale-check clean
