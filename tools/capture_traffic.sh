#!/bin/bash
# Run on the GPU box: the C4 bench line, then an `ncu --set full` capture of the dominant kernel family at exactly the
# shape the line's roofline was timed on; writes profiles-ready text + the JSON bench.py reads for roofline.traffic.
#   bash tools/capture_traffic.sh <config> <out-prefix>
cfg=${1:-c4}; out=${2:-gpurun_out/r02}
python bench.py --config $cfg --steps 10 --warmup 3 > ${out}_bench_${cfg}.json 2> ${out}_bench_${cfg}.err || { tail -5 ${out}_bench_${cfg}.err; exit 1; }
read net kind shape <<< $(python - <<PY
import json
d=json.load(open("${out}_bench_${cfg}.json"))["roofline"]
net,kind=d["kernel"].split(":")[:2]
print(net, kind.strip(), json.dumps(d["timed_shape"]).replace(" ",""))
PY
)
case $kind in Conv) rx="conv_tc|pwconv|conv_simt|stem_conv";; DwConv) rx="dwconv";; CtcHead) rx="ctc_head";; *) rx="$kind";; esac
echo "dominant family $net:$kind at $shape -> ncu -k regex:$rx"
ncu --set full --clock-control none --profile-from-start off -k regex:"$rx" -o /tmp/fam -f python tools/ncu_workload.py $net "$shape" > ${out}_ncu_${cfg}.log 2>&1
python tools/ncu_traffic.py /tmp/fam.ncu-rep "$rx" > ${out}_ncu_${cfg}_family.txt
tail -2 ${out}_ncu_${cfg}_family.txt
python - <<PY
import json,re
t=open("${out}_ncu_${cfg}_family.txt").read()
j=json.loads(re.search(r"# json (.*)",t).group(1))
json.dump({"$cfg":{"family":"$net:$kind","shape":json.loads("$shape"),"launches":j["launches"],"traffic_bytes_per_launch":j["traffic_bytes_per_launch"],
  "source":"profiles/r02_ncu_${cfg}_family.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum over the family's launches of one forward pass)"}},
  open("${out}_traffic_${cfg}.json","w"),indent=1)
PY
