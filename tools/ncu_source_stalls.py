"""Per-kernel stall mix and hottest CUDA source lines of an `ncu --set full --import-source on` report (warp-state samples,
`ncu --page source --print-source cuda,sass --csv`).  usage: python tools/ncu_source_stalls.py <report.ncu-rep> <kernel-id>...
where kernel-id is ncu's `::regex:<name>:<n-th launch>`; prints markdown."""
import csv
import io
import subprocess
import sys


def load(rep, kid):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", kid],
                         capture_output=True, text=True).stdout
    cur, hdr, lines, name = None, None, [], ""
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            name = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue                                              # SASS rows of the mixed view
        # a source line that contains quotes breaks the CSV fields on the left: index from the right
        off = len(r) - len(hdr)
        get = lambda n: int(r[hdr.index(n) + off] or 0)           # noqa: E731
        st = {n[6:]: get(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
        text = ",".join(r[1:2 + off]).strip()
        lines.append((cur, int(r[0]), get("# Samples"), get("Instructions Executed"), text, st))
    return name, lines


def main():
    rep = sys.argv[1]
    for kid in sys.argv[2:]:
        name, lines = load(rep, kid)
        tot = sum(l[2] for l in lines) or 1
        mix = {}
        for l in lines:
            for k, v in l[5].items():
                mix[k] = mix.get(k, 0) + v
        short = name.rsplit("(", 1)[0].replace("void kws::<unnamed>::", "").replace("(bool)", "").replace("(int)", "")
        print(f"### `{short}` ({kid})\n")
        print(f"{tot} warp-state samples, {sum(l[3] for l in lines) / 1e6:.0f} M warp instructions.  Stall mix: " +
              ", ".join(f"{k} {100 * v / tot:.0f} %" for k, v in sorted(mix.items(), key=lambda x: -x[1])[:7]) + "\n")
        print("| file:line | samples | share | instructions | top stalls | source |")
        print("|---|---|---|---|---|---|")
        for l in sorted(lines, key=lambda l: -l[2])[:12]:
            top = ", ".join(f"{k} {v}" for k, v in sorted(l[5].items(), key=lambda x: -x[1])[:2] if v)
            src = l[4].replace("|", "\\|")[:90]
            print(f"| {l[0]}:{l[1]} | {l[2]} | {100 * l[2] / tot:.1f} % | {l[3]} | {top} | `{src}` |")
        print()


if __name__ == "__main__":
    main()
