"""DRAM traffic per kernel from an `ncu --set full` report -> profiles/ncu_traffic_<tag>.json (read by bench.py for
roofline.traffic). Usage: python tools/ncu_traffic.py rep.ncu-rep out.json "<source note>" samples_per_launch"""
import csv
import json
import subprocess
import sys

STAGES = {"gate": "k_gate_partial", "yin_fp32": "k_yin_corr", "yin_decide": "k_yin_decide_reg", "marks": "k_marks",
          "voc_autocorr": "k_voc_autocorr2", "voc_levinson": "k_voc_levinson_static", "voc_synth": "k_voc_synth_stream",
          "pitch_lpc": "k_pitch_autocorr", "pitch_psola": "k_pitch_psola", "pitch_iir": "k_pitch_iir", "mix": "k_mix"}


def tobytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(rep, out, note, samples):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = {"source": note, "samples_per_launch_capture": int(samples), "stages": {}}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        for stage, key in STAGES.items():
            if key in name and stage not in res["stages"]:
                b = sum(tobytes(r[col[k]], units[col[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                res["stages"][stage] = {"kernel": name, "dram_bytes_per_launch_capture": b, "dram_bytes_per_sample": b / float(samples)}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res["stages"], indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
