"""Rebuilds profiles/r02_ncu_traffic.json (what bench.py's roofline.traffic reads) from `ncu --page raw --csv` exports, so the
figure is derived, not typed:
    python tools/make_ncu_traffic.py <gemm_shapes.csv> <attention.csv> <gemm_shapes_timing.json> [out.json] [attention_b8.csv]
attention_b8.csv: ncu --set full of the bench-shape forward (batch 8, bounded scores: AFB_DIAG_BOUND=12 tools/diag_attn_time.py)
gemm_shapes.csv : ncu --set full -k regex:gemm_bf16 -c 9 over tools/profile_gemm_shapes.py --once (launch i = entry i)
attention.csv   : ncu --set full -k regex:attention --launch-skip 12 --launch-count 3 over tools/diag_attn_bwd_time.py
                  (forward, dQ, dK/dV in that order)
Also writes a per-launch table (time, DRAM read / write, algorithmic bytes, tensor-pipe share) next to it."""
import csv
import json
import sys

ORDER = ["img_qkv", "img_attn_out", "img_mlp_up", "img_mlp_down", "single_qkv", "single_mlp_up", "single_proj_out", "lora_a_up",
         "img_qkv_fused_norm_rope"]
WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "time",
        "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_hmma",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "launch__registers_per_thread": "regs",
        "smsp__cycles_active.avg": "cycles", "sm__cycles_elapsed.max": "cycles_elapsed"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
        "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def read(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    out = []
    for r in rows[hdr_i + 2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m, key in WANT.items():
            if m in hdr:
                j = hdr.index(m)
                try:
                    v = float(r[j].replace(",", ""))
                except ValueError:
                    continue
                d[key] = v * UNIT.get(units[j], 1.0)
        out.append(d)
    return out


def main():
    gemm_csv, attn_csv, timing_json = sys.argv[1:4]
    out_path = sys.argv[4] if len(sys.argv) > 4 else "profiles/r02_ncu_traffic.json"
    attn_b8 = read(sys.argv[5])[0] if len(sys.argv) > 5 else None
    timing = {l["name"]: l for l in json.load(open(timing_json))["launches"]}
    g = read(gemm_csv)
    table = []
    for name, d in zip(ORDER, g):
        t = timing[name]
        table.append(dict(name=name, desc=t["desc"], ms_ncu=d["time"] * 1e3, ms_events=t["ms"], tflops_events=t["tflops"],
                          dram_read_bytes=d.get("dram_read"), dram_write_bytes=d.get("dram_write"),
                          algorithmic_bytes=t["algorithmic_bytes"],
                          traffic_over_algorithmic=(d.get("dram_read", 0) + d.get("dram_write", 0)) / t["algorithmic_bytes"],
                          tensor_pipe_pct=d.get("tensor_pct", d.get("tensor_pct_hmma")), l2_hit_pct=d.get("l2_hit_pct"),
                          registers=d.get("regs")))
    a = read(attn_csv)
    top = next(r for r in table if r["name"] == "single_mlp_up")
    rec = dict(source=f"derived by tools/make_ncu_traffic.py from {gemm_csv} and {attn_csv} (ncu --set full --clock-control none, "
                      "final round-2 kernels; per-launch values, cold cache)",
               gemm=dict(kernel="gemm_bf16_2cta_kernel", launch=top["desc"] + " (the largest launch of the step)",
                         dram_bytes_per_launch=top["dram_read_bytes"] + top["dram_write_bytes"],
                         dram_read_bytes=top["dram_read_bytes"], dram_write_bytes=top["dram_write_bytes"],
                         algorithmic_bytes_per_launch=top["algorithmic_bytes"], ms_ncu=top["ms_ncu"]),
               gemm_launch_table=table,
               attention=(dict(kernel=attn_b8["kernel"].split("(")[0], launch="batch 8 x 24 heads x S 4608, bounded scores (the bench shape)",
                               dram_bytes_per_launch=attn_b8.get("dram_read", 0) + attn_b8.get("dram_write", 0),
                               algorithmic_bytes_per_launch=4 * 8 * 4608 * 3072 * 2, ms_ncu=attn_b8["time"] * 1e3,
                               tensor_pipe_pct=attn_b8.get("tensor_pct", attn_b8.get("tensor_pct_hmma")), registers=attn_b8.get("regs"))
                          if attn_b8 else None),
               attention_training_shape=dict(kernel=a[0]["kernel"].split("(")[0],
                                             launch="batch 4 x 24 heads x S 4608, " + ("bounded scores" if "split" in a[0]["kernel"] else "running-max kernel (no score bound passed by the timing tool)"),
                                             dram_bytes_per_launch=a[0].get("dram_read", 0) + a[0].get("dram_write", 0),
                                             algorithmic_bytes_per_launch=4 * 4 * 4608 * 3072 * 2, ms_ncu=a[0]["time"] * 1e3,
                                             tensor_pipe_pct=a[0].get("tensor_pct", a[0].get("tensor_pct_hmma"))),
               attention_backward=[dict(kernel=r["kernel"].split("(")[0], ms_ncu=r["time"] * 1e3,
                                        dram_bytes_per_launch=r.get("dram_read", 0) + r.get("dram_write", 0),
                                        tensor_pipe_pct=r.get("tensor_pct", r.get("tensor_pct_hmma")), registers=r.get("regs"))
                                   for r in a[1:]])
    json.dump(rec, open(out_path, "w"), indent=1)
    print(json.dumps(rec, indent=1)[:3000])


if __name__ == "__main__":
    main()
