# Multi-GPU capture: the GPU suite on a box with N GPUs (the tests that need >= 2 devices run here), then bench.py under
# torchrun exactly as the driver launches it.  Usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh <tag> N'
cd $GRAFT_REPO_ROOT
TAG=${1:-r2_n}; N=${2:-2}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/gpu.txt 2>&1
nvidia-smi topo -m >> $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_full_size_gpu.py::test_c4_all_256_pairs_against_compiled_reference tests/test_parity_gpu.py -q -m gpu -k "multi_gpu or pool or c4 or shard" 2>&1 | tail -6 | tee $O/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"; cat $O/bench_n$N.json; tail -3 $O/bench_n$N.err
