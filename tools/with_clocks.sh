nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/clk_$1.csv &
P=$!
shift
"$@"
kill $P
