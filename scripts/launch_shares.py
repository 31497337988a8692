"""Shares of serialised kernel time by kernel name from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X`): python scripts/launch_shares.py X.csv[.gz]"""
import collections, csv, gzip, io, re, sys


def shares(path):
    raw = gzip.open(path, "rt", errors="ignore").read() if path.endswith(".gz") else open(path, errors="ignore").read()
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"[(<].*", "", name).replace("dmt::", "")
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[mu], 1e-3)
        tot[name] += v
        cnt[name] += 1
    return tot, cnt


if __name__ == "__main__":
    tot, cnt = shares(sys.argv[1])
    S = sum(tot.values())
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in tot.most_common(30):
        print("| {} | {} | {:.2f} | {:.1f} % |".format(k[:64], cnt[k], v / cnt[k], 100 * v / S))
    print("total {:.2f} ms over {} launches".format(S / 1e3, sum(cnt.values())))
