"""Distil profiles/traffic.json (DRAM bytes of the dominant conv launch, read by bench.py) from an `ncu --set full` report whose
first captured kernel is downs.0.0.block1 at the given batch.  usage: ncu_traffic.py prof_conv.ncu-rep 1024 <source note>"""
import csv, json, subprocess, sys
rep, batch, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, first = rows[0], rows[1], rows[2]
ix = {h: i for i, h in enumerate(hdr)}
def val(name):
    v, u = float(first[ix[name]].replace(",", "")), units[ix[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
out = {"downs.0.0.block1": {"batch": batch, "kernel": first[ix["Kernel Name"]][:40], "dram_read": val("dram__bytes_read.sum"),
                            "dram_write": val("dram__bytes_write.sum"), "source": note}}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
