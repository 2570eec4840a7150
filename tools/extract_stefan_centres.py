"""Derive the 40 per-channel centres used by `--center stefan` from the
reference's measurement log (data/Stefan_file.txt, read by
src/GPPupilDemodulation.jl:84-104: the lines starting with `avg`, fields 2, 3
and 5 after whitespace splitting = channel name, VX [mV], VY [mV]).

Only the 40 averaged values are kept (decimal strings verbatim so that the
decimal->double conversion is the reference's).  Run in the build container,
where /root/reference exists; the output is committed.
"""
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/Stefan_file.txt"
dst = sys.argv[2] if len(sys.argv) > 2 else "gppupildemodulation.jl_b200/data/stefan_centres.txt"
rows = []
with open(src) as fh:
    for line in fh:
        if line.startswith("avg"):
            v = line.split()
            rows.append((v[1], v[2], v[4]))
assert len(rows) == 40, len(rows)
with open(dst, "w") as out:
    out.write("# channel  VX[mV]  VY[mV]   (40 `avg` rows of the reference's data/Stefan_file.txt)\n")
    for name, vx, vy in rows:
        out.write(f"{name} {vx} {vy}\n")
print("wrote", dst)
