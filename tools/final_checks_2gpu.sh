#!/bin/bash
# sanity on 2 GPUs over NCCL: the (rows x particles) grid against the single-GPU run, the row-sharded selector, the README example
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/check_row_sharding.py --grid 2x1 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/check_row_sharding.py --grid 1x2 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/check_sharded_selector.py 2>&1 | tail -2
python examples/readme_regression.py 2>&1 | tail -3
