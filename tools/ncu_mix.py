"""Executed instruction mix and DRAM bytes per stage from an ncu metrics CSV -> profiles/ncu_mix_<tag>.json (read by
bench.py for roofline.mix and roofline.traffic).

    ncu --metrics <METRICS below> --clock-control none --csv --log-file X.csv python bench.py --steps 1 --warmup 1 ...
    python tools/ncu_mix.py out.json "<note>" workload=X.csv:samples_per_pass [workload2=Y.csv:samples ...]

Every launch in the capture is attributed to a stage; counts are divided by the number of passes captured (= launches of
k_mix) and by the samples of one pass: executed thread-instructions per audio sample."""
import collections
import csv
import json
import sys

METRICS = ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,"
           "smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum")

# kernel name prefix -> stage of bench.py's stage_ms (first match wins)
STAGE_OF = [("k_gate", "gate"), ("k_yin_corr", "yin_fp32"), ("k_yin_decide", "yin_decide"), ("k_yin<", "yin_fp64_recheck"),
            ("k_yin", "yin_fp64_recheck"), ("k_marks", "marks"), ("k_voc_autocorr", "voc_autocorr"), ("k_voc_levinson", "voc_levinson"),
            ("k_voc_gain", "voc_levinson"), ("k_voc_synth", "voc_synth"), ("k_voc_tail", "voc_synth"), ("k_pitch_autocorr", "pitch_lpc"),
            ("k_pitch_levinson", "pitch_lpc"), ("k_pitch_psola", "pitch_psola"), ("k_pitch_iir", "pitch_iir"), ("k_mix", "mix"),
            ("k_carry", "other"), ("k_hist", "other")]
SHORT = {"smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd",
         "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": "ffma",
         "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum": "fadd", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum": "fmul",
         "smsp__inst_executed.sum": "warp_inst", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
         "gpu__time_duration.sum": "ns"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3,
        "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def stage_of(kernel):
    k = kernel.replace("void ", "")
    for prefix, st in STAGE_OF:
        if k.startswith(prefix):
            return st
    return None  # not an engine kernel (input generator, issue-rate microbenchmarks)


def parse(path, samples):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    stages = collections.OrderedDict()
    launches = collections.Counter()
    seen = set()
    for r in rows[1:]:
        st = stage_of(r[ik])
        if st is None or r[im] not in SHORT:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= UNIT.get(r[iu], 1.0)
        d = stages.setdefault(st, collections.Counter())
        d[SHORT[r[im]]] += v
        if (r[iid], r[ik]) not in seen:
            seen.add((r[iid], r[ik]))
            launches[r[ik].split("(")[0]] += 1
    passes = max(sum(n for k, n in launches.items() if stage_of(k) == "mix"), 1)  # one mix kernel (k_mix / k_mix4<..>) per pass
    out = collections.OrderedDict()
    for st, d in stages.items():
        per = {k: v / passes / float(samples) for k, v in d.items()}
        out[st] = {"fp64_ops_per_sample": per.get("dfma", 0) + per.get("dadd", 0) + per.get("dmul", 0),
                   "fp32_ops_per_sample": per.get("ffma", 0) + per.get("fadd", 0) + per.get("fmul", 0),
                   "dfma_per_sample": per.get("dfma", 0), "ffma_per_sample": per.get("ffma", 0),
                   "thread_inst_per_sample": 32.0 * per.get("warp_inst", 0),
                   "dram_bytes_per_sample": per.get("dram_read", 0) + per.get("dram_write", 0),
                   "ms_per_pass_under_ncu": d.get("ns", 0) / passes * 1e-6}
    tot = {k: sum(s[k] for s in out.values()) for k in ("fp64_ops_per_sample", "fp32_ops_per_sample", "dram_bytes_per_sample", "thread_inst_per_sample")}
    return {"samples_per_pass_capture": int(samples), "passes_captured": passes, "launches": dict(launches), "stages": out, "total": tot}


def main(out, note, *specs):
    res = {"source": note, "metrics": METRICS, "workloads": {}}
    for sp in specs:
        name, rest = sp.split("=", 1)
        path, samples = rest.rsplit(":", 1)
        res["workloads"][name] = parse(path, float(samples))
        print(name, json.dumps(res["workloads"][name]["total"]))
    with open(out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
