"""Per-launch hot spots of a `ncu --set full --import-source on` capture: the most-sampled SASS instructions (warp stall
sampling, all samples) with their execution counts -- the view that showed the store epilogue, not the gathers, bounding
the shared-MLP layers.

    python tools/hot_sass.py gpurun_out/prof_mlp_r02c.ncu-rep 33 profiles/ncu_mlp_r02c_hot_sass.md
"""
import csv, io, subprocess, sys

rep, n_launch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = [f"# hottest SASS instructions per launch of `{rep.split('/')[-1]}` (warp-stall samples, `ncu --page source`)\n",
         "Reading guide: a `BRA` with an execution count of (tiles x 4) / (tiles x 8.. x chunks) / (tiles) is the try-wait loop of the "
         "epilogue warps (accumulator full) / the producer warps (stage empty) / the MMA warp (stage full); `@P0 EXIT` = warps that "
         "finished early (tail); arithmetic right after a load (`FADD`, `FMUL`, `IMAD.MOV`) = the wait for that load.\n"]
for L in range(n_launch):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(L),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    name = rows[0][1] if len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ia, isrc, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) > iex and r[ia].isdigit()]
    data = data[:len(data) // 2] if len(data) > 2 and data[0][isrc] == data[len(data) // 2][isrc] else data
    tot = sum(int(r[ia]) for r in data) or 1
    top = sorted(data, key=lambda r: -int(r[ia]))[:8]
    short = name.split("mlp_layer_kernel")[-1].split("(pvn3d")[0]
    lines.append(f"\n## launch {L}: mlp_layer_kernel{short} -- {tot} samples\n")
    lines.append("| share | executed | instruction |\n|---:|---:|---|")
    for r in top:
        lines.append(f"| {100 * int(r[ia]) / tot:.1f}% | {r[iex]} | `{r[isrc].strip()[:70]}` |")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(lines))
