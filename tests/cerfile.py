"""Test tooling: a plain-Python reader of isONclust2 `.cer` batches used to inspect what the
reference binary and the isonclust2_b200 CLI write.  Not product code."""
import struct


class _R:
    def __init__(self, d):
        self.d, self.o = d, 0

    def rd(self, f):
        v = struct.unpack_from("<" + f, self.d, self.o)
        self.o += struct.calcsize("<" + f)
        return v[0]

    def s(self):
        n = self.rd("Q")
        v = self.d[self.o:self.o + n]
        self.o += n
        return v

    def skip(self, n):
        self.o += n


def _seq(r):
    if not r.rd("B"):
        return None
    return dict(name=r.s(), seq=r.s(), qual=r.s(), score=r.rd("d"), err=r.rd("d"))


def _mins(r):
    n = r.rd("Q")
    v = r.d[r.o:r.o + 12 * n]
    r.o += 12 * n
    return v


def _graph(r):
    o0 = r.o
    r.rd("I")
    for elt in (4, 4):
        r.skip(r.rd("Q") * elt)
    for _ in range(r.rd("Q")):
        if r.rd("B"):
            r.skip(8)
    for _ in range(r.rd("Q")):
        if r.rd("B"):
            r.skip(r.rd("Q") * 4)
            r.skip(8)
    for elt in (4, 8, 8, 4, 4):
        r.skip(r.rd("Q") * elt)
    return r.d[o0:r.o]


def parse(path):
    """Returns dict(header fields, args, mindb (list of (key, tuple(cls)) in file order), mindb_span, cls, graphs)."""
    d = open(path, "rb").read()
    r = _R(d)
    b = dict(BatchNr=r.rd("i"), BatchStart=r.rd("Q"), BatchEnd=r.rd("Q"), BatchBases=r.rd("Q"), TotalReads=r.rd("i"), NrCls=r.rd("i"))
    a = dict(Verbose=r.rd("B"), Debug=r.rd("B"), InFastq=r.s())
    for n in ("KmerSize", "BatchSize", "BatchMaxSeq", "WindowSize", "MinShared", "ConsMinSize", "ConsMaxSize", "ConsPeriod", "MinClsSize"):
        a[n] = r.rd("i")
    for n in ("MinQual", "MappedThreshold", "AlignedThreshold", "MinFraction", "MinProbNoHits"):
        a[n] = r.rd("d")
    a["BatchOutFolder"] = r.s()
    a["Mode"] = r.rd("i")
    b["args"] = a
    b.update(LeftLeaf=r.s(), RightLeaf=r.s(), Depth=r.rd("i"))
    m0 = r.o
    mindb = []
    for _ in range(r.rd("Q")):
        key = r.rd("I")
        m = r.rd("Q")
        mindb.append((key, struct.unpack_from("<%dI" % m, d, r.o)))
        r.skip(4 * m)
    b["mindb"], b["mindb_span"] = mindb, (m0, r.o)
    cls = []
    for _ in range(r.rd("Q")):
        cid = r.rd("I")
        if cid == 0:
            cls.append(None)
            continue
        assert cid & 0x80000000, "aliased cluster pointer"
        members = []
        for _ in range(r.rd("Q")):
            pid = r.rd("I")
            if pid == 0:
                members.append(None)
                continue
            assert pid & 0x80000000, "aliased read pointer"
            members.append(dict(raw=_seq(r), hpc=_seq(r), mins=_mins(r), rev=_mins(r), strand=r.rd("i"), id=r.s()))
        cls.append(members)
    b["cls"] = cls
    graphs = []
    for _ in range(r.rd("Q")):
        graphs.append(_graph(r) if r.rd("B") else None)
    b["graphs"] = graphs
    assert r.o == len(d), "trailing bytes"
    b["bytes"] = d
    return b


def same_up_to_mindb_order(pa, pb):
    """True if two parsed batches are byte-identical outside the MinDB section and hold the same MinDB entries."""
    (a0, a1), (b0, b1) = pa["mindb_span"], pb["mindb_span"]
    return (pa["bytes"][:a0] == pb["bytes"][:b0] and pa["bytes"][a1:] == pb["bytes"][b1:] and
            sorted(pa["mindb"]) == sorted(pb["mindb"]))


def merge_decisions(left, right, out):
    """The per-right-cluster decisions (best, strand) the reference took in `cluster -l left -r right -o out`, derived
    from cluster membership in its output: best = -2 skipped, -1 founded a new cluster, else the left cluster joined."""
    n_left = len(left["cls"])
    where = {}
    for ci, members in enumerate(out["cls"]):
        for m in members:
            where.setdefault(m["id"], []).append((ci, m["strand"]))
    dec = []
    for members in right["cls"]:
        if not members:
            dec.append((-2, 0))
            continue
        probe = members[1] if len(members) > 1 else members[0]
        hits = where.get(probe["id"])
        if not hits:
            dec.append((-2, 0))
            continue
        ci, st = hits[-1]
        if ci >= n_left:
            dec.append((-1, 0))
        else:
            dec.append((ci, 1 if st == probe["strand"] else -1))
    return dec


def same_file_except_total_reads(a, b):
    """Byte equality of two `.cer` files with Batch::TotalReads (bytes 28..31) masked: the reference's `sort` never
    initialises that field (serialize.h:29), `cluster` copies whatever the sorted batch held into its
    output, and nothing ever reads it — so two runs of the REFERENCE differ there too; this build writes 0."""
    da, db = open(a, "rb").read(), open(b, "rb").read()
    return len(da) == len(db) and da[:28] == db[:28] and da[32:] == db[32:]
