#!/bin/bash
# usage: clocks_during.sh <command...>  -- samples SM clock / power / power-cap flag every 50 ms while the command runs
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu --format=csv,noheader,nounits -i 0 -lms 50 > /tmp/clk.csv &
SMI=$!
"$@"
kill $SMI
python3 - <<'PY'
rows=[l.strip().split(', ') for l in open('/tmp/clk.csv') if l.strip()]
busy=[r for r in rows if float(r[1])>400]
import statistics as st
if busy:
    print("samples under load %d: sm clock median %.0f MHz (min %.0f max %.0f), power median %.0f W, power-cap active in %d%% of samples, temp %s C"%(
        len(busy), st.median(float(r[0]) for r in busy), min(float(r[0]) for r in busy), max(float(r[0]) for r in busy),
        st.median(float(r[1]) for r in busy), 100*sum(r[2].startswith('Active') for r in busy)//len(busy), busy[-1][4]))
else:
    print("no samples under load", rows[:3])
PY
