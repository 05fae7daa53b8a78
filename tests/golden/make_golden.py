#!/usr/bin/env python3
"""Builds tests/golden/*.json from the reference's own `call -m` test cases (test/test.pl:276-308).

Run in the build container only (needs /root/reference); the fixtures it writes are committed because
/root/reference does not exist on the GPU box.  Nothing is computed here: the script only TRANSCODES
  * the input VCF of each case into the flat arrays that cross the C-ABI (what bcf_get_format_int32 /
    bcf_get_info_float hand to mcall(), mcall.c:1444-1510), applying the driver logic that sits in front
    of mcall(): -S sample subsetting and sexes (vcfcall.c:270-344, PED 202-261), --ploidy-file and the
    per-record ploidy state machine (ploidy.c:192-230, vcfcall.c:807-825), the unseen-allele detection and
    the `-v` pre-filter (vcfcall.c:1101-1115), -G group files (mcall.c:297-348);
  * the expected `.out` VCF of the case into per-record expectations (ALT, QUAL text, AC, AN, GT, PL, GQ, GP).

usage: python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
TEST = os.path.join(REF, "test")
OUT = os.path.dirname(os.path.abspath(__file__))

MISSING = -2**31
VEC_END = -2**31 + 1

# (fixture name, input vcf, expected out, dict of options)  -- test/test.pl:276-308, `-m` cases without -C alleles / -g
CASES = [
    ("mpileup.1", "mpileup.vcf", "mpileup.1.out", dict(v=1)),
    ("mpileup.3", "mpileup.vcf", "mpileup.3.out", dict(v=1, S="mpileup.3.samples")),
    ("mpileup.4", "mpileup.vcf", "mpileup.4.out", dict(v=1, S="mpileup.4.samples")),
    ("mpileup.5", "mpileup.vcf", "mpileup.5.out", dict(v=1, S="mpileup.5.samples")),
    ("mpileup.X", "mpileup.X.vcf", "mpileup.X.out", dict(v=1, ploidy="mpileup.ploidy", S="mpileup.samples")),
    ("mpileup.X.ped", "mpileup.X.vcf", "mpileup.X.out", dict(v=1, ploidy="mpileup.ploidy", S="mpileup.ped")),
    ("mpileup.X.2", "mpileup.X.vcf", "mpileup.X.2.out", dict(v=1, ploidy="mpileup.ploidy", S="mpileup.2.samples")),
    ("mpileup.hwe.1", "mpileup.NA19213.NA19129.vcf", "mpileup.hwe.1.out", dict(v=1)),
    ("mpileup.hwe.1b", "mpileup.NA19213.NA19129.vcf", "mpileup.hwe.1b.out", dict(v=1, G="-", Gtag="AD")),
    ("mpileup.hwe.2", "mpileup.hwe.vcf", "mpileup.hwe.2.out", dict(v=1)),
    ("mpileup.hwe.3", "mpileup.hwe.vcf", "mpileup.hwe.3.out", dict(v=1, G="-", Gtag="AD")),
    ("mpileup.hwe.4", "mpileup.hwe.vcf", "mpileup.hwe.4.out", dict(v=1, G="mpileup.hwe.samples", Gtag="AD")),
    ("call-G.1", "call-G.vcf", "call-G.1.out", dict(v=1)),
    ("call-G.2", "call-G.vcf", "call-G.2.out", dict(v=1, G="-", Gtag="AD")),
    ("call-G.2.1", "call-G.2.vcf", "call-G.2.1.out", dict(v=1, F=("AN_POP", "AC_POP"))),
    ("call.af-fixation.1", "call.af-fixation.vcf", "call.af-fixation.1.out", dict()),
    ("call.af-fixation.2", "call.af-fixation.vcf", "call.af-fixation.2.out", dict(G="call.af-fixation.txt")),
    ("call.af-fixation.3", "call.af-fixation.vcf", "call.af-fixation.3.out", dict(G="call.af-fixation.txt", a="GP,GQ")),
]


def read_vcf(path):
    samples, recs = [], []
    fmt_defs = set()
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith("##"):
            if line.startswith("##FORMAT=<ID="):
                fmt_defs.add(line.split("ID=")[1].split(",")[0])
            continue
        f = line.split("\t")
        if line.startswith("#"):
            samples = f[9:]
            continue
        info = {}
        for kv in f[7].split(";"):
            if "=" in kv:
                k, v = kv.split("=", 1)
                info[k] = v
        fmt = f[8].split(":") if len(f) > 8 else []
        smpl = [dict(zip(fmt, s.split(":"))) for s in f[9:]]
        recs.append(dict(chrom=f[0], pos=int(f[1]), ref=f[3], alt=[] if f[4] == "." else f[4].split(","),
                         qual=f[5], info=info, fmt=fmt, smpl=smpl))
    return samples, recs, fmt_defs


def int_vector(txt, n):
    """One sample's integer FORMAT vector as bcf_get_format_int32 returns it, padded to n with vector_end [htslib]."""
    if txt is None or txt == ".":
        return [MISSING] + [VEC_END] * (n - 1)
    v = [MISSING if x == "." else int(x) for x in txt.split(",")]
    assert len(v) <= n, (txt, n)
    return v + [VEC_END] * (n - len(v))


def parse_samples_file(path):
    """vcfcall.c:270-344 incl. the PED form (202-261).  Returns [(name, sex-or-ploidy string)]."""
    lines = [l.rstrip("\n") for l in open(path) if l.strip()]
    ped = []
    for l in lines:
        c = l.split()
        if len(c) < 6:      # the PED parser needs 5 separators, i.e. a 6th column start
            ped = None
            break
        ped.append((c[1], "M" if c[4][0] == "1" else "F"))
    if ped is not None:
        return ped
    out = []
    for l in lines:
        if l.lstrip().startswith("#"):
            continue
        c = l.split()
        out.append((c[0], c[1] if len(c) > 1 else "2"))
    return out


def parse_ploidy(path):
    regions, dflt = [], {}
    sexes = []
    for l in open(path):
        c = l.split()
        if not c:
            continue
        if c[3] not in sexes:
            sexes.append(c[3])
        if c[0] == "*":
            dflt[c[3]] = int(c[4])
        else:
            regions.append((c[0], int(c[1]), int(c[2]), c[3], int(c[4])))
    return dict(regions=regions, dflt=dflt, sexes=sexes, global_dflt=2)


def ploidy_query(pl, chrom, pos):
    """ploidy.c:192-230: returns {sex: ploidy}."""
    hits = [r for r in pl["regions"] if r[0] == chrom and r[1] <= pos <= r[2]]
    if not hits:
        return {s: pl["dflt"].get(s, pl["global_dflt"]) for s in pl["sexes"]}
    res = {s: pl["global_dflt"] for s in pl["sexes"]}
    for r in hits:
        if r[4] != pl["global_dflt"]:
            res[r[3]] = r[4]
    return res


def build_case(name, vcf, out, opt):
    samples, recs, fmt_defs = read_vcf(os.path.join(TEST, vcf))
    # ---- samples / sexes (vcfcall.c:270-344)
    if "S" in opt:
        sel = parse_samples_file(os.path.join(TEST, opt["S"]))
        sel = [(n, s) for n, s in sel if n in samples]
    else:
        sel = [(n, None) for n in samples]
    idx = [samples.index(n) for n, _ in sel]
    names = [n for n, _ in sel]
    S = len(names)
    # ---- ploidy (vcfcall.c:1068-1072 default: all diploid)
    ploidy_def = parse_ploidy(os.path.join(TEST, opt["ploidy"])) if "ploidy" in opt else None
    if ploidy_def:
        for _, s in sel:        # ploidy_add_sex for sexes only named in the samples file
            if s is not None and s not in ("0", "1", "2") and s not in ploidy_def["sexes"]:
                ploidy_def["sexes"].append(s)
    ploidy_vectors = [[2] * S]
    cur_ploidy = [2] * S                      # aux.ploidy = ploidy_max at init (vcfcall.c:652-655)
    prev_sex2 = None
    if ploidy_def:
        pmax = max([2] + [r[4] for r in ploidy_def["regions"]] + list(ploidy_def["dflt"].values()))
        cur_ploidy = [pmax] * S
        ploidy_vectors = [list(cur_ploidy)]
        prev_sex2 = {s: pmax for s in ploidy_def["sexes"]}
    # ---- groups (mcall.c:250-349)
    groups = None
    if "G" in opt:
        if opt["G"] == "-":
            groups = [[i] for i in range(S)]
        else:
            order, members = [], {}
            for l in open(os.path.join(TEST, opt["G"])):
                c = l.split()
                if len(c) < 2 or c[0] not in names:
                    continue
                if c[1] not in members:
                    members[c[1]] = []
                    order.append(c[1])
                members[c[1]].append(names.index(c[0]))
            groups = [sorted(members[g]) for g in order]
            assert sorted(sum(groups, [])) == list(range(S))
    gtag = opt.get("Gtag")
    if groups and not gtag:
        gtag = "QS" if "QS" in fmt_defs else "AD"     # mcall.c:272-281
    flag = 2 if opt.get("v") else 0
    tags = 0
    for t in opt.get("a", "").split(","):
        tags |= {"GQ": 64, "GP": 128, "": 0}[t]

    sites = []
    for r in recs:
        als = [r["ref"]] + r["alt"]
        A = len(als)
        unseen = 0
        for i in range(1, A):               # vcfcall.c:1101-1111
            a = als[i]
            if a[0] == "X" or a[:3] in ("<X>", "<*>"):
                unseen = i
                break
        is_ref = A == 1 or (A == 2 and unseen > 0)
        if is_ref and (flag & 2):
            continue
        if ploidy_def:                      # set_ploidy, vcfcall.c:807-825
            sex2 = ploidy_query(ploidy_def, r["chrom"], r["pos"])
            if sex2 != prev_sex2:
                cur_ploidy = [(int(s) if s in ("0", "1", "2") else sex2.get(s if s is not None else ploidy_def["sexes"][-1]))
                              for _, s in sel]
                prev_sex2 = sex2
        if cur_ploidy not in ploidy_vectors:
            ploidy_vectors.append(list(cur_ploidy))
        G = A * (A + 1) // 2
        pl = [int_vector(r["smpl"][i].get("PL"), G) for i in idx]
        site = dict(chrom=r["chrom"], pos=r["pos"], alleles=als, unseen=unseen, ploidy_id=ploidy_vectors.index(cur_ploidy), pl=pl)
        if "QS" in r["info"]:
            site["qs"] = r["info"]["QS"].split(",")       # kept as text: float32 parse happens in the test
        if groups:
            site["ad"] = [int_vector(r["smpl"][i].get(gtag), A) for i in idx]
        if "F" in opt:
            an, ac = opt["F"]
            if an in r["info"]:
                site["prior_an"] = int(r["info"][an])
            if ac in r["info"]:
                site["prior_ac"] = [MISSING if x == "." else int(x) for x in r["info"][ac].split(",")]
        sites.append(site)

    # ---- expectations
    osamples, orecs, _ = read_vcf(os.path.join(TEST, out))
    assert osamples == names, (name, osamples, names)
    expect = []
    for r in orecs:
        e = dict(chrom=r["chrom"], pos=r["pos"], alleles=[r["ref"]] + r["alt"], qual=r["qual"],
                 ac=[int(x) for x in r["info"]["AC"].split(",")] if "AC" in r["info"] else [],
                 an=int(r["info"]["AN"]), gt=[s["GT"] for s in r["smpl"]])
        if "PL" in r["fmt"]:
            e["pl"] = [s["PL"] for s in r["smpl"]]
        if "GQ" in r["fmt"]:
            e["gq"] = [s["GQ"] for s in r["smpl"]]
        if "GP" in r["fmt"]:
            e["gp"] = [s["GP"] for s in r["smpl"]]
        expect.append(e)
    return dict(name=name, source=dict(vcf="test/" + vcf, out="test/" + out, options={k: v for k, v in opt.items()}),
                nsmpl=S, samples=names, flag=flag, output_tags=tags, theta=1.1e-3, groups=groups,
                use_prior="F" in opt, ploidy_vectors=ploidy_vectors, sites=sites, expect=expect)


def main():
    for name, vcf, out, opt in CASES:
        case = build_case(name, vcf, out, opt)
        path = os.path.join(OUT, name + ".json")
        with open(path, "w") as fh:
            json.dump(case, fh, separators=(",", ":"))
        print(f"{name}: {len(case['sites'])} sites into mcall, {len(case['expect'])} expected records, "
              f"{os.path.getsize(path)//1024} kB")


if __name__ == "__main__":
    main()
