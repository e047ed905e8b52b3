for mb in 48 96 144 192 288; do
  PP_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /tmp/cs_$mb.json 2>/dev/null
  python - $mb <<'PY'
import sys, json
mb = sys.argv[1]
for l in open("/tmp/cs_%s.json" % mb):
    if l.startswith("{"):
        d = json.loads(l); print("chunk MB", mb, "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 1), "device", round(d["ms_per_step"], 1))
PY
done
