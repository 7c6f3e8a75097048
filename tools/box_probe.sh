set -x
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread|Core"
python -c "import numpy as np; np.show_runtime()" 2>&1 | grep -A20 simd | head -24
python tools/tie_probe.py
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
ncu --query-metrics 2>/dev/null | grep -i -E "dmma|fp64|pipe_tensor" | head -60 > gpurun_out/ncu_metrics_dmma.txt
wc -l gpurun_out/ncu_metrics_dmma.txt
