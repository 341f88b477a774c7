# Convenience targets; everything is also reachable through python directly (README.md).
PY ?= python

.PHONY: build oracle test test-gpu bench bench-reference clean

build:            ## libclawb200.so for sm_100a (nvcc cross-compiles without a GPU)
	$(PY) -m pyclaw_b200.build

oracle:           ## CPU oracle (test infrastructure only)
	$(MAKE) -C oracle

test: build oracle ## CPU suite: oracle vs goldens, host logic, C ABI, 2-rank gloo
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build oracle ## parity suite on a B200
	$(PY) -m pytest tests -x -q -m gpu

bench: build      ## one JSON line, Euler 8192^2 on one GPU
	$(PY) bench.py

bench-reference: oracle ## the CPU arm
	$(PY) bench.py --impl reference

clean:
	rm -f pyclaw_b200/csrc/*.so oracle/*.so
