true
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 12 -c 6 --csv --log-file gpurun_out/l2hint_traffic.csv python tools/profile_step.py 3 > /dev/null 2>&1
ILCC2_LIB=build_variants/lib_nohint.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -s 12 -c 6 --csv --log-file gpurun_out/l2nohint_traffic.csv python tools/profile_step.py 3 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("l2nohint","l2hint"):
    rows=[r for r in csv.reader(l for l in open(f"gpurun_out/{f}_traffic.csv") if l.startswith('"'))]
    h=rows[0]; tot=0
    for r in rows[1:]:
        if "dram__bytes" in r[h.index("Metric Name")]:
            v=float(r[h.index("Metric Value")].replace(",","")); u=r[h.index("Metric Unit")]
            tot+=v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
    print(f, "DRAM MB per step", round(tot/1e6,1), "ratio", round(tot/472305664,3))
PY
