for v in 1 0; do
echo "QD_PDL=$v"
QD_PDL=$v python bench.py --no-e2e --no-cpu --no-checks > gpurun_out/ab_$v.json 2>/dev/null
python - <<P
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("headline", round(d["roofline"]["frac"],4))
for k,v in d["extra"].items():
    if isinstance(v,dict) and "roofline" in v: print(" ",k, round(v["roofline"]["frac"],4), round(v.get("ms",0),4))
P
done
