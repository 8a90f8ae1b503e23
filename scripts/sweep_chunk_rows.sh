for r in 20 80 160; do
  FRINGE_CHUNK_ROWS=$r timeout 250 python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | grep "^{" | tail -1 > gpurun_out/sweep_$r.json
done
