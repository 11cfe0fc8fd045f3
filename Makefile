# Convenience targets; everything here is also documented in README.md.
PY ?= python

.PHONY: build test-cpu test-gpu bench bench-reference smoke clean

build:            ## nvcc (sm_100a) -> nlp_b200/liblda_b200.so, gcc -> the CPU oracles
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build   ## oracle vs golden vectors, C-ABI exports, host logic, gloo world-size-2 (no GPU needed)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## parity of the CUDA paths against the oracles (needs a B200)
	$(PY) -m pytest tests -q -m gpu

smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## BASELINE.json's metric, one JSON line
	$(PY) bench.py --steps 5 --warmup 3

bench-reference:  ## the reference's CPU algorithm (C restatement) on the host cores
	$(PY) bench.py --impl reference --steps 5 --warmup 3

clean:
	rm -f nlp_b200/liblda_b200.so oracle/liblda_oracle.so oracle/libindex_oracle.so tests/cpp/lda_test tests/cpp/index_test tools/micro/fp64_peak
