"""Build profiles/ncu_traffic.json from `ncu --set full` reports: per bench key, dram__bytes_read.sum +
dram__bytes_write.sum of ONE launch, with the capture file and the git hash of the build that was profiled.
usage: python scripts/ncu_traffic.py <git-hash> key=report.ncu-rep[:kernel-substring[:index]] ..."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(rep):
    if rep.endswith(".csv"):                      # `ncu -i report --page raw --csv` taken on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        def val(name):
            v, u = float(r[ix[name]].replace(',', '')), units[ix[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        res.append({"kernel": r[ix["Kernel Name"]], "bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "us": float(r[ix["gpu__time_duration.sum"]].replace(',', '')) *
                    {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}[units[ix["gpu__time_duration.sum"]]]})
    return res


def main():
    git = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    for spec in sys.argv[2:]:
        key, rest = spec.split("=", 1)
        parts = rest.split(":")
        rep, sub, idx = parts[0], (parts[1] if len(parts) > 1 else ""), (int(parts[2]) if len(parts) > 2 else 0)
        ls = [l for l in launches(rep) if sub in l["kernel"]]
        l = ls[idx]
        d[key] = {"bytes": int(l["bytes"]), "capture": os.path.basename(rep), "git": git, "kernel": l["kernel"][:80],
                  "us_under_ncu": round(l["us"], 2)}
        print(key, d[key])
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
