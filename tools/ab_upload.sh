# A/B of the split-upload settings (XFG_UPLOAD_STREAMS, XFG_UPLOAD_SPLIT) on the e2e latency; run under gpurun from the repo root
for cfg in "1:1,1,1,1,1,1,1" "2:1,1,1,1,1,1,1" "3:1,1,1,1,1,1,1" "4:1,1,1,1,1,1,1" "3:1,1,2,3"; do
  XFG_UPLOAD_STREAMS=${cfg%%:*} XFG_UPLOAD_SPLIT=${cfg##*:} python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${cfg%%:*}_$(echo ${cfg##*:} | tr -d ,).json 2>/dev/null
done
