"""Read an `ncu --page raw --csv` export WITH its units row and print / return per-kernel DRAM bytes and duration.

    python profiles/ncu_units.py profiles/r01d_prof_gain_f4s_raw.csv [...]
    python profiles/ncu_units.py --rebuild-traffic      # rewrites profiles/ncu_traffic.json from the captures it names

ncu scales every column separately (the round-1 table added "42.165 Gbyte" and "21.478 Mbyte" as if both were GB);
this parser multiplies each value by ITS unit."""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": d.get("Kernel Name", "")}
        for key, name in (("dram__bytes_read.sum", "dram_read_bytes"), ("dram__bytes_write.sum", "dram_write_bytes"),
                          ("gpu__time_duration.sum", "seconds")):
            if key in d:
                try:
                    rec[name] = float(d[key].replace(",", "")) * SCALE[units[hdr.index(key)]]
                except (ValueError, KeyError):
                    rec[name] = None
        out.append(rec)
    return out


TRAFFIC_SOURCES = {
    "c4:f4:1": ("r02_prof_gain_f4s_raw.csv", "gemm_f4s_2sm_kernel<EPI_GAIN>, 480189 x 17770, full pass = first launch of the capture"),
    "c5:product:1": ("r02za_prof_c5_raw.csv", "bool_product_panel_list_kernel<dynamic>, 1M x 100k, k = 64", "bool_product"),
    "c5:confusion:1": ("r02za_prof_c5_raw.csv", "confusion_panel_list_kernel<|gt| known, count mode 3>, 1M x 100k, k = 64", "confusion"),
    "c4:i8:1": ("r01b_prof_gain_raw.csv", "gemm_i8_2sm_kernel<EPI_GAIN>"),
    "c4:pq-i8:1": ("r01b_prof_gain2_raw.csv", "gemm_i8_2sm_kernel<EPI_GAIN2>, w_fp = 0.2"),
    "c4:pq-f4:1": ("r01e_prof_gain2_f4_raw.csv", "gemm_f4_2sm_kernel<EPI_GAIN2>, w_fp = 0.2"),
}


def rebuild_traffic(extra=None):
    table = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, each value scaled by "
                         "ITS unit (profiles/ncu_units.py --rebuild-traffic); bench.py copies the matching entry into "
                         "roofline.traffic (workload:kernel-variant:gpus)"}
    src = dict(TRAFFIC_SOURCES)
    src.update(extra or {})
    for key, spec in src.items():
        fn, what = spec[0], spec[1]
        name = spec[2] if len(spec) > 2 else ""                 # substring of the kernel name when a capture holds several
        path = os.path.join(HERE, fn)
        if not os.path.exists(path):
            continue
        rec = [r for r in parse(path) if r.get("dram_read_bytes") is not None and name in r["kernel"]][0]
        table[key] = {"bytes": rec["dram_read_bytes"] + rec["dram_write_bytes"],
                      "read_bytes": rec["dram_read_bytes"], "write_bytes": rec["dram_write_bytes"],
                      "kernel_seconds_under_ncu": rec.get("seconds"),
                      "source": "profiles/%s (%s)" % (fn, what)}
    json.dump(table, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)
    return table


if __name__ == "__main__":
    if "--rebuild-traffic" in sys.argv:
        print(json.dumps(rebuild_traffic(), indent=1))
    else:
        for p in sys.argv[1:]:
            for rec in parse(p):
                print(p, json.dumps(rec))
