python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 >> gpurun_out/final_tests.log
for v in t14 t13 t12; do
  if [ $v = t14 ]; then unset XFG_LIB; else export XFG_LIB=$PWD/ab/$v.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  timeout 300 python bench.py --n-log2 16 --ext 1 --steps 20 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/ab16_$v.json 2> gpurun_out/ab16_$v.err
done
cat gpurun_out/final_tests.log
