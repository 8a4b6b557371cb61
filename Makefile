# Convenience targets; the driver uses __graft_entry__.build() / smoke() and bench.py directly.
PY ?= python

build:            ## nvcc -gencode arch=compute_100a,code=sm_100a -> shippingenv_b200/libshipenv.so, gcc -> oracle/_build
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU suite: oracle vs reference fixtures, ABI, host logic, gloo
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build   ## CUDA vs fixtures / oracle, bit-exact (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench: build      ## one JSON line: env-steps/s, roofline, e2e, cpu_baseline
	$(PY) bench.py

golden:           ## regenerate tests/golden/*.npz from the unmodified reference ($$SHIPPINGENV_REF or /root/reference)
	PYTHONDONTWRITEBYTECODE=1 $(PY) tests/golden/make_golden.py

.PHONY: build test test-gpu bench golden
