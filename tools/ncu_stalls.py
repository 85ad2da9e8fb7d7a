"""Summarise an ncu report's source page: python tools/ncu_stalls.py report.ncu-rep [top]
Per SASS instruction: stall samples with the two dominant reasons; mbarrier wait sites with the barrier offset;
instruction mix.  (ncu -i ... --page source --csv needs -lineinfo builds and --import-source on captures.)"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    S = lambda r: int(r[ix["# Samples"]])
    tot = sum(S(r) for r in data)
    print("total samples", tot, " instructions", len(data))
    for r in sorted(data, key=lambda r: -S(r))[:top_n]:
        st = sorted([(int(r[ix[h]]), h[6:]) for h in stalls], reverse=True)[:2]
        print(f"{S(r):7d} {100 * S(r) / tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {st}")
    print("-- wait sites")
    for i, r in enumerate(data):
        if "NANOSLEEP" in r[ix["Source"]] and S(r) > tot * 0.002:
            j = i
            while j > 0 and "PHASECHK" not in data[j][ix["Source"]]:
                j -= 1
            print(f"{S(r):7d} {100 * S(r) / tot:5.1f}%  {data[j][ix['Source']].strip()[:80]}")
    print("-- stall totals", sorted(((sum(int(r[ix[h]]) for r in data), h[6:]) for h in stalls), reverse=True)[:8])
    ex = lambda key: sum(int(r[ix["Instructions Executed"]]) for r in data if key in r[ix["Source"]])
    print("-- warp instructions executed: total", sum(int(r[ix["Instructions Executed"]]) for r in data),
          {k: ex(k) for k in ("MUFU", "FFMA2", "FMNMX", "F2FP", "FADD2", "LDTM", "STTM", "UTCHMMA", "UTCIMMA", "LDL", "STL", "BRA", "SYNCS", "SHFL")})


if __name__ == "__main__":
    main()
