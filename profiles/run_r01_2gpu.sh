# 2-GPU check (gpurun --gpus 2): tests + smoke on GPU 0, then the bench under torchrun for c2 and the c4 training step.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -k 5 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -k 5 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_n1.json 2>/dev/null; python profiles/summarize_bench.py < gpurun_out/bench_c2_n1.json | head -2
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -2 gpurun_out/bench_c2_n2.err; python profiles/summarize_bench.py < gpurun_out/bench_c2_n2.json | head -2
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; python profiles/summarize_bench.py < gpurun_out/bench_c4_n2.json | head -2
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
