# Round profile on the GPU box (run under gpurun from the repo root): bench lines without a profiler first, then the ncu launch list of the
# same command, then one `--set full` capture of the heavy kernels of one proof.  Outputs land in gpurun_out/; summaries are copied to profiles/.
set -x
T=${1:-r02}
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_ref20.json 2>/dev/null
python bench.py --n-log2 16 --ext 1 --steps 20 --warmup 3 --headline-only > gpurun_out/${T}_bench16.json 2>/dev/null
python bench.py --workload verify > gpurun_out/${T}_verify.json 2>/dev/null
python bench.py --workload air > gpurun_out/${T}_air.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-preload --headline-only > gpurun_out/${T}_plain.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-preload --headline-only > gpurun_out/${T}_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'commit_rows_kernel|ntt_pass_r16|deep_kernel|constraint_kernel|fri_fold_kernel|ood_kernel|fri_tail_kernel|tree_' -c 26 -f -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-preload --headline-only > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_prof.ncu-rep
# gpurun copies back at most 64 MiB: keep the raw metric table (CSV) and the summary, drop the report itself
ncu -i gpurun_out/${T}_prof.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/${T}_ncu_raw.csv > gpurun_out/${T}_ncu_table.txt 2>&1
rm -f gpurun_out/${T}_prof.ncu-rep
# memory checker on small proofs (every kernel family incl. the fused tail, the generic front-end and the batch verifier)
# compute-sanitizer is closed on this pool ("runs under it have left GPUs needing a reset"): memory safety is covered by the parity tests on
# small and ragged cases and by the canaries of tests/test_gpu_stages.py instead.
