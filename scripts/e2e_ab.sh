for cfg in "BWT_COMPUTE_TOKEN=1 D=2" "BWT_COMPUTE_TOKEN=0 D=2" "BWT_COMPUTE_TOKEN=0 D=3" "BWT_COMPUTE_TOKEN=0 D=4"; do
  D=$(echo $cfg | sed 's/.*D=//'); E=$(echo $cfg | cut -d' ' -f1)
  echo "$cfg: $(env $E timeout -s KILL 120 python bench.py --cpu-seconds 0 --steps 3 --warmup 2 --depth $D 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("value %.1fM e2e %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6))')"
done
