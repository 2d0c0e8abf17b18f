# Convenience targets (the driver's own entry points are __graft_entry__.build / smoke and bench.py).
PY ?= python

build:            ## compile the CUDA library for sm_100a (cross-compiles without a GPU)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## oracle vs golden vectors, ABI, host logic (no GPU)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke:
	$(PY) __graft_entry__.py smoke

bench:            ## BASELINE config 2 on one GPU; `make bench N=8` under torchrun
ifeq ($(N),)
	$(PY) bench.py
else
	$(PY) -m torch.distributed.run --nnodes=1 --nproc-per-node $(N) --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $(N)
endif

golden-check:     ## re-extract tests/golden/reference_specs.json from /root/reference and compare
	$(PY) tests/golden/extract_reference_specs.py

.PHONY: build test-cpu test-gpu smoke bench golden-check
