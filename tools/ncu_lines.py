"""Per-source-line instruction / shared-memory wavefront / stall-sample totals of one kernel from an .ncu-rep (read here).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <pairs> [file-substring] [min-inst-per-pair] [min-wavefronts-per-pair]
`pairs` normalises the counts (e.g. N*K); only lines of source files whose path contains the substring are listed."""
import csv
import io
import subprocess
import sys


def main():
    rep, pairs = sys.argv[1], float(sys.argv[2])
    sub = sys.argv[3] if len(sys.argv) > 3 else 'local_step_fast'
    mi = float(sys.argv[4]) if len(sys.argv) > 4 else 3.0
    mw = float(sys.argv[5]) if len(sys.argv) > 5 else 1.5
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True,
                         text=True).stdout
    cur, hdr, ix, agg = None, None, {}, []
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == 'File Path':
            cur = r[1]
            continue
        if r and r[0] == 'Line No':
            hdr, ix = r, {}
            for i, h in enumerate(r):
                ix.setdefault(h, i)
            continue
        if hdr is None or len(r) < len(hdr) or not cur or sub not in cur or r[2] != '-':
            continue
        try:
            e = float(r[ix['Instructions Executed']] or 0)
            w = float(r[ix['L1 Wavefronts Shared']] or 0)
            s = float(r[ix['# Samples']] or 0)
        except ValueError:
            continue
        agg.append((int(r[0]), e / pairs, w / pairs, s, r[1]))
    tots = sum(a[3] for a in agg) or 1.0
    print('total: inst/pair %.1f  wavefronts/pair %.1f' % (sum(a[1] for a in agg), sum(a[2] for a in agg)))
    for a in agg:
        if a[1] > mi or a[2] > mw:
            print('%4d inst/pair %7.1f wf/pair %7.1f samp%% %5.1f | %s' % (a[0], a[1], a[2], 100 * a[3] / tots, a[4][:100]))


if __name__ == '__main__':
    main()
