nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 100 > /tmp/clk.csv &
SMI=$!
python tools/biggrid.py 65536 65536
python tools/freerun.py | head -3
kill $SMI
sort /tmp/clk.csv | uniq -c | sort -rn | head -12
