"""Second, independent restatement of the hot path, written straight from the Rust sources.

TEST INFRASTRUCTURE ONLY (tests/test_oracle_crosscheck.py).  It shares no code with oracle/lcr_oracle.cpp
and does not include include/lcr_contract.h: every function below was transcribed from the cited lines of
/root/reference/src/*.rs, with Python floats (IEEE f64, libm log10 / pow as Rust's std) and plain loops in
the reference's order.  A transcription error in either restatement shows up as a disagreement between
the two; an error of the numerical contract (fixed point, lcr_log10 ...) shows up against this file's libm
arithmetic.  What it covers:

    R0  isolated regions               util.rs:236-332
    P0  read filter + fetch window     util.rs:636-668
    P1  pileup                         util.rs:669-949
    P2  two major alleles              util.rs:158-176
    P3-P7 candidate cascade, genotype likelihood, routing, dense filters   candidate.rs:24-51, 54-528
    F1/F3 fragments, link counts       fragment.rs:10-309
    S0-S5 aki, the two sweeps, cross_optimize, the 2^n enumeration of phase()   phase.rs:32-49, 77-176, 257-355, 673-691, 810-976, 1087-1122

The only piece that is not in the reference is the random source: the contract replaces thread_rng by the
counter-based generator documented in include/lcr_contract.h (splitmix64 finaliser); it is restated here
from that description.
"""
import math

MAX_BASE_QUALITY = 30  # main.rs:19-21
M64 = (1 << 64) - 1


# ---------------------------------------------------------------- random source (contract, not reference)
def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def region_key(tid, start):
    return mix64(((tid & 0xFFFFFFFF) << 32) | start)


def uniform(seed, rkey, stream, call, idx):
    h = mix64(seed ^ rkey)
    h = mix64(h ^ ((stream << 32) | call))
    h = mix64(h ^ idx)
    return (h >> 11) * (1.0 / 9007199254740992.0)


RNG_INIT_SIGMA = 1


# ---------------------------------------------------------------- small helpers
def f32(x):
    import struct

    return struct.unpack("f", struct.pack("f", x))[0]


def leading_softclips(cig):
    """rust-htslib CigarStringView::leading_softclips: the soft clip at the start, looking through one hard clip."""
    if not cig:
        return 0
    if cig[0][0] == "S":
        return cig[0][1]
    if cig[0][0] == "H" and len(cig) > 1 and cig[1][0] == "S":
        return cig[1][1]
    return 0


def trailing_softclips(cig):
    if not cig:
        return 0
    if cig[-1][0] == "S":
        return cig[-1][1]
    if cig[-1][0] == "H" and len(cig) > 1 and cig[-2][0] == "S":
        return cig[-2][1]
    return 0


def read_passes(P, r):
    """util.rs:652-668 (the same test opens fragment.rs:28-49)"""
    if r["mapq"] < P["min_mapq"] or len(r["seq"]) < P["min_read_length"] or r["flag"] & 0x4 or r["flag"] & 0x100 or r["flag"] & 0x800:
        return False
    de = r["de"]
    if de == de and f32(de) >= f32(P["divergence"]):  # Aux::Float only; NaN stands for "no de:f tag"
        return False
    return True


def in_fetch_window(region, r):
    """bam.fetch((chr, start, end)) with the region's 1-based numbers handed to htslib's 0-based half-open query"""
    rlen = sum(n for op, n in r["cigar"] if op in "MDN=X")
    endpos = r["pos"] + (rlen if rlen else 1)
    return r["pos"] < region["end"] and endpos > region["start"]


# ---------------------------------------------------------------- P1 pileup (util.rs:621-949)
def new_basefreq(ref_base):
    return dict(a=0, c=0, g=0, t=0, n=0, d=0, ref_base=ref_base, bq=dict(A=[], C=[], G=[], T=[]),
                strands=dict(A=[0, 0], C=[0, 0], G=[0, 0], T=[0, 0]), ts=[0, 0])


class BadCigar(Exception):
    pass


def pileup(P, region, reads, ref_seq):
    vec_size = region["end"] - region["start"]
    start0 = region["start"] - 1
    fv = [new_basefreq(chr(ref_seq[start0 + i])) for i in range(vec_size)]
    polya = P["polya_tail_length"]
    n_pass = n_aligned = 0
    for r in reads:
        if not in_fetch_window(region, r) or not read_passes(P, r):
            continue
        n_pass += 1
        seq, qual, cig = r["seq"], r["qual"], r["cigar"]
        strand = 1 if r["flag"] & 0x10 else 0
        ts = r["ts"]
        lead, trail = leading_softclips(cig), trailing_softclips(cig)
        pos_fv = r["pos"] - start0
        pos_read = lead if lead > 0 else 0
        stop = False
        for op, n in cig:
            if stop:
                break
            if op in "SH":
                continue
            if op in "MX=":
                for _ in range(n):
                    if pos_fv < 0:
                        pos_fv += 1
                        pos_read += 1
                        continue
                    if pos_fv >= vec_size:
                        break
                    n_aligned += 1
                    if pos_read >= len(seq):
                        raise BadCigar()
                    base = chr(seq[pos_read])
                    baseq = qual[pos_read] if qual[pos_read] < MAX_BASE_QUALITY else MAX_BASE_QUALITY
                    bf = fv[pos_fv]
                    ref_base = bf["ref_base"]
                    poly_a = homopolymer = trim = False
                    dist_end = P["distance_to_read_end"]
                    curr = pos_read
                    read_end_boundary = len(seq) - trail
                    near = abs(curr - lead) < dist_end or abs(curr - read_end_boundary) < dist_end
                    if P["platform"] == 1 and near:  # Platform::Ont: trim the read ends
                        trim = True
                    if not trim and near:
                        for tmpi in range(curr - polya, curr + 2):
                            if tmpi < 0 or tmpi + polya - 1 >= len(seq):
                                continue
                            pc = [0, 0, 0, 0]  # A, T, C, G
                            for tmpj in range(polya):
                                b = chr(seq[tmpi + tmpj])
                                if b == "A" and ref_base != "A":
                                    pc[0] += 1
                                elif b == "T" and ref_base != "T":
                                    pc[1] += 1
                                elif b == "C" and ref_base != "C":
                                    pc[2] += 1
                                elif b == "G" and ref_base != "G":
                                    pc[3] += 1
                            if pc[0] >= polya or pc[1] >= polya:
                                poly_a = True
                            if pc[2] >= polya or pc[3] >= polya:
                                homopolymer = True
                    if not trim and not poly_a and not homopolymer:
                        if strand == 0:
                            if ts == "+":
                                bf["ts"][0] += 1
                            elif ts == "-":
                                bf["ts"][1] += 1
                        else:
                            if ts == "+":
                                bf["ts"][1] += 1
                            elif ts == "-":
                                bf["ts"][0] += 1
                        u = base.upper()
                        if u in "ACGT" and base in "ACGTacgt":
                            bf[u.lower()] += 1
                            bf["bq"][u].append(baseq)
                            bf["strands"][u][strand] += 1
                    pos_fv += 1
                    pos_read += 1
            elif op == "D":
                for _ in range(n):
                    if pos_fv < 0:
                        pos_fv += 1
                        continue
                    if pos_fv >= vec_size:
                        break
                    fv[pos_fv]["d"] += 1
                    pos_fv += 1
            elif op == "I":
                if pos_fv < 1:
                    pos_read += n
                    continue
                if pos_fv >= vec_size:
                    stop = True  # `break` leaves the loop over the CIGAR (util.rs:924-926)
                    break
                pos_read += n
            elif op == "N":
                for _ in range(n):
                    if pos_fv < 0:
                        pos_fv += 1
                        continue
                    if pos_fv >= vec_size:
                        break
                    fv[pos_fv]["n"] += 1
                    pos_fv += 1
            else:
                raise BadCigar()  # util.rs:943-945 panics
    return fv, n_pass, n_aligned


# ---------------------------------------------------------------- P2-P7 candidates (candidate.rs:54-528)
def two_major(bf):
    """util.rs:162-176; Python's sort is stable like Rust's sort_by"""
    x = sorted([("A", bf["a"]), ("C", bf["c"]), ("G", bf["g"]), ("T", bf["t"])], key=lambda t: -t[1])
    rb = bf["ref_base"]
    if x[0][0] != rb and x[1][0] != rb:
        if x[2][1] == x[1][1] and x[2][0] == rb:
            return x[0][0], x[0][1], x[2][0], x[2][1]
        if x[3][1] == x[1][1] and x[3][0] == rb:
            return x[0][0], x[0][1], x[3][0], x[3][1]
    return x[0][0], x[0][1], x[1][0], x[1][1]


def strand_odds_ratio(ref_fw, ref_rv, alt_fw, alt_rv):
    """candidate.rs:24-35, every operation rounded to f32"""
    x00, x01, x10, x11 = f32(ref_fw + 1), f32(ref_rv + 1), f32(alt_fw + 1), f32(alt_rv + 1)
    sym = f32(f32(f32(x00 * x11) / f32(x01 * x10)) + f32(f32(x01 * x10) / f32(x00 * x11)))
    ref_ratio = f32(min(x00, x01) / max(x00, x01))
    alt_ratio = f32(min(x10, x11) / max(x10, x11))
    ln = lambda v: f32(math.log(v))
    return f32(f32(ln(sym) + ln(ref_ratio)) - ln(alt_ratio))


def binomial_two_tailed(k, n):
    """candidate.rs:37-47 with the exact binomial(n, 1/2) distribution (statrs evaluates the same numbers through beta_reg)"""
    from fractions import Fraction

    cdf = lambda j: sum(Fraction(math.comb(n, i), 2 ** n) for i in range(0, j + 1))
    if k == 0:
        return 2 * cdf(0)
    if k == n:
        return 2 * (1 - cdf(n - 1))
    return 2 * min(cdf(k), 1 - cdf(k - 1))


SOR_THRESHOLD = strand_odds_ratio(5, 5, 9, 1)


def candidates(P, region, fv, exons=None):
    """exons: None, or the (start, stop) intervals of parse_annotation (util.rs:435-439) for --exon-only (candidate.rs:80-89)."""
    cands, homo, het, edit, somatic = [], [], [], [], []
    position = region["start"] - 1
    for bf in fv:
        pos = position
        position += 1
        if exons is not None and not any(s < pos + 2 and e > pos + 1 for s, e in exons):  # Lapper::find(position + 1, position + 2)
            continue
        total = bf["a"] + bf["c"] + bf["g"] + bf["t"]
        if total < P["min_depth"] or total > P["max_depth"]:
            continue
        a1, c1, a2, c2 = two_major(bf)
        f1, f2 = f32(f32(c1) / f32(total)), f32(f32(c2) / f32(total))
        rb = bf["ref_base"]
        if a1 == rb:
            ref_allele, alt = a1, [(a2, f2, c2)]
        elif a2 == rb:
            ref_allele, alt = a2, [(a1, f1, c1)]
        else:
            ref_allele, alt = rb, [(a1, f1, c1), (a2, f2, c2)]
        if ref_allele not in "ACGTacgt":  # VALID_ALLELES, main.rs:23
            continue
        if len(alt) == 1:
            if total < 200 and alt[0][1] < f32(P["low_allele_frac_cutoff"]):
                continue
            if total >= 200 and alt[0][2] < P["low_allele_cnt_cutoff"]:
                continue
        if bf["d"] >= alt[0][2]:
            continue
        depth_incl = total + bf["d"] + bf["n"]
        if f32(f32(c1 + c2) / f32(depth_incl)) < f32(P["min_allele_freq_include_intron"]):
            continue
        if a1 != rb:
            if c1 > 0 and sum(1 for q in bf["bq"][a1] if q >= P["min_baseq"]) < 2:
                continue
        elif a2 != rb:
            if c2 > 0 and sum(1 for q in bf["bq"][a2] if q >= P["min_baseq"]) < 2:
                continue
        if P["use_strand_bias"]:
            rs = bf["strands"].get(ref_allele.upper(), [0, 0]) if ref_allele in "ACGT" else [0, 0]
            sors = [strand_odds_ratio(rs[0], rs[1], *bf["strands"][b]) for b, _, _ in alt]
            if max(sors) > SOR_THRESHOLD:
                continue
            if len(alt) == 1:
                fw, rv = bf["strands"][alt[0][0]]
                if fw + rv <= 30 and binomial_two_tailed(fw, fw + rv) < 0.05:
                    continue
                if fw * rv == 0:
                    continue
        if rb not in "ACGT":  # candidate.rs:243-264: N or an unknown (e.g. lower-case) reference byte
            continue
        # candidate.rs:267-282: the identical-base qualities first, then the three other letters in A, C, G, T order
        ll = [0.0, 0.0, 0.0]
        for q in bf["bq"][rb]:
            e = math.pow(0.1, q / 10.0)
            ll[0] += math.log10(e) if e > 0 else -math.inf
            ll[2] += math.log10(1.0 - e) if 1.0 - e > 0 else -math.inf
        for b in "ACGT":
            if b == rb:
                continue
            for q in bf["bq"][b]:
                e = math.pow(0.1, q / 10.0)
                ll[0] += math.log10(1.0 - e) if 1.0 - e > 0 else -math.inf
                ll[2] += math.log10(e) if e > 0 else -math.inf
        ll[1] -= total * math.log10(2.0)
        theta = 0.001
        bg = [theta / 2.0, theta, 1.0 - 1.5 * theta]
        lp = [ll[i] + math.log10(bg[i]) for i in range(3)]
        mx = max(lp)
        lp = [v - mx for v in lp]
        vp = [math.pow(10.0, v) for v in lp]
        s = sum(vp)
        vp = [v / s for v in vp]
        variant_quality = -10.0 * math.log10(max(10e-301, vp[2]))
        mxl = max(ll)
        l10 = [math.pow(10.0, v - mxl) for v in ll]
        s = l10[0] + l10[1] + l10[2]
        gp = [v / s for v in l10]
        ph = sorted(-10.0 * (math.log10(v) if v > 0 else -math.inf) for v in gp)
        genotype_quality = ph[1] - ph[0]
        if gp[0] > gp[1] and gp[0] > gp[2]:
            vt, gt = 2, -1
        elif gp[1] > gp[0] and gp[1] > gp[2]:
            vt, gt = 1, 0
        else:
            vt, gt = 0, 1
        c = dict(pos=pos, alleles=(a1, a2), freqs=(f1, f2), reference=rb, depth=total, variant_quality=variant_quality, gp=gp,
                 genotype_quality=genotype_quality, variant_type=vt, genotype=gt, haplotype=0, rna_editing=False, for_phasing=False,
                 cand_somatic=False, hom_var=False, het_var=False, dense=False, cover=[])
        if variant_quality < P["min_qual"]:
            continue
        fwd_ts, rev_ts = bf["ts"]
        if ref_allele == "A" and alt[0][0] == "G" and (fwd_ts > rev_ts * 2 or (fwd_ts == 0 and rev_ts == 0)) and vt != 2:
            c["rna_editing"] = True
            cands.append(c)
            edit.append(len(cands) - 1)
            continue
        if ref_allele == "T" and alt[0][0] == "C" and (rev_ts > fwd_ts * 2 or (fwd_ts == 0 and rev_ts == 0)) and vt != 2:
            c["rna_editing"] = True
            cands.append(c)
            edit.append(len(cands) - 1)
            continue
        if len(alt) == 1 and alt[0][1] < f32(P["min_allele_freq"]):
            c["cand_somatic"] = True
            cands.append(c)
            somatic.append(len(cands) - 1)
            continue
        if vt == 2:
            if len(alt) == 2 and alt[0][1] >= f32(P["min_allele_freq"]) and alt[1][1] >= f32(P["min_allele_freq"]):
                c["variant_type"], c["genotype"] = 3, -1
            c["hom_var"] = c["for_phasing"] = True
            cands.append(c)
            homo.append(len(cands) - 1)
            continue
        if vt == 1:
            if len(alt) == 2:
                c["variant_type"], c["genotype"] = 3, -1
                c["hom_var"] = c["for_phasing"] = True
                cands.append(c)
                homo.append(len(cands) - 1)
            else:
                c["het_var"] = c["for_phasing"] = True
                cands.append(c)
                het.append(len(cands) - 1)
            continue
    idx = sorted(homo + het)
    for win, ge, mincnt in ((P["dense_win_size"], False, P["min_dense_cnt"]), (5, True, 3)):
        for i in range(len(idx)):
            sp = cands[idx[i]]["pos"]
            for j in range(i, len(idx)):
                diff = cands[idx[j]]["pos"] - sp
                if (diff >= win) if ge else (diff > win):
                    if j - i >= mincnt:
                        for tk in range(i, j):
                            cands[idx[tk]]["dense"] = True
                            cands[idx[tk]]["for_phasing"] = False
                    break
                if j == len(idx) - 1 and j - i + 1 >= mincnt:
                    for tk in range(i, j):
                        cands[idx[tk]]["dense"] = True
                        cands[idx[tk]]["for_phasing"] = False
    return cands, edit, somatic


# ---------------------------------------------------------------- F1 / F3 fragments (fragment.rs:10-309)
def fragments(P, region, reads, cands):
    frags = []
    if not cands:
        return frags
    for ridx, r in enumerate(reads):
        if not in_fetch_window(region, r) or not read_passes(P, r):
            continue
        pos = r["pos"]
        if pos > cands[-1]["pos"]:
            continue
        seq, qual, cig = r["seq"], r["qual"], r["cigar"]
        pos_ref, pos_q = pos, leading_softclips(cig)
        idx = 0
        if not pos <= cands[0]["pos"]:
            while idx < len(cands) and cands[idx]["pos"] < pos:
                idx += 1
        snp_pos = cands[idx]["pos"]
        frag = dict(read=ridx, list=[], haplotag=0, for_phasing=False, links=0)
        for op, n in cig:
            if op in "SH":
                continue
            if op in "MX=":
                for _ in range(n):
                    if pos_ref == snp_pos:
                        if pos_q >= len(seq):
                            raise BadCigar()
                        c = cands[idx]
                        base = chr(seq[pos_q])
                        baseq = qual[pos_q] if qual[pos_q] < 30 else 30
                        if base == c["reference"]:
                            p = 1
                        elif base in c["alleles"] and base != c["reference"]:
                            p = -1
                        else:
                            p = 0
                        if not c["dense"] and p != 0:
                            frag["list"].append(dict(snp=idx, base=base, baseq=baseq, p=p, prob=math.pow(10.0, -baseq / 10.0), phase_site=c["for_phasing"]))
                        idx += 1
                        if idx < len(cands):
                            snp_pos = cands[idx]["pos"]
                    pos_q += 1
                    pos_ref += 1
            elif op == "I":
                pos_q += n
            elif op in "DN":
                for _ in range(n):
                    if pos_ref == snp_pos:
                        idx += 1
                        if idx < len(cands):
                            snp_pos = cands[idx]["pos"]
                    pos_ref += 1
            else:
                raise BadCigar()
        frag["links"] = sum(1 for fe in frag["list"] if fe["phase_site"])
        frag["for_phasing"] = frag["links"] >= P["min_linkers"]
        for fe in frag["list"]:
            cands[fe["snp"]]["cover"].append(len(frags))
        frags.append(frag)
    return frags


# ---------------------------------------------------------------- S0-S5 phasing (phase.rs)
def aki(sigma, delta, eta, p, prob):
    x = sigma * delta if eta == 0 else eta
    return 1.0 - prob if p == x else prob


def log10(v):
    return math.log10(v) if v > 0 else (-math.inf if v == 0 else math.nan)


def cal_sigma_delta_eta_log(sigma_k, delta, eta, ps, probs):
    q1 = q2 = q3 = 0.0
    for i in range(len(delta)):
        q1 += log10(aki(sigma_k, delta[i], eta[i], ps[i], probs[i]))
    for i in range(len(delta)):
        q2 += log10(aki(1, delta[i], eta[i], ps[i], probs[i]))
        q3 += log10(aki(-1, delta[i], eta[i], ps[i], probs[i]))
    return 1.0 - q1 / (q2 + q3)


def cal_delta_eta_sigma_log(delta_i, eta_i, sigma, ps, probs):
    prior_homref = math.log10(1.0 - 1.5 * 0.001)
    prior_homvar = math.log10(0.5 * 0.001)
    cov = len(sigma)
    prior_het = math.log10(0.001) if cov == 0 else math.log10(0.001) - cov * math.log10(2.0)
    q1 = 0.0
    for k in range(cov):
        q1 += log10(aki(sigma[k], delta_i, eta_i, ps[k], probs[k]))
    q1 += prior_het if eta_i == 0 else (prior_homref if eta_i == 1 else prior_homvar)
    q2 = q3 = q4 = q5 = 0.0
    for k in range(cov):
        q2 += log10(aki(sigma[k], delta_i, -1, ps[k], probs[k]))
        q3 += log10(aki(sigma[k], delta_i, 0, ps[k], probs[k]))
        q4 += log10(aki(sigma[k], delta_i, 1, ps[k], probs[k]))
        q5 += log10(aki(sigma[k], -delta_i, 0, ps[k], probs[k]))
    q2 += prior_homvar
    q3 += prior_het
    q4 += prior_homref
    q5 += prior_het
    return 1.0 - q1 / (q2 + q3 + q4 + q5)


def overall_probability(cands, frags):
    logp = 0.0
    for f in frags:
        if not f["for_phasing"] or f.get("ds_skip") or f["haplotag"] == 0:  # ds_skip: apply_downsampling && !downsampled (phase.rs:262)
            continue
        for fe in f["list"]:
            if fe["phase_site"]:
                c = cands[fe["snp"]]
                logp += log10(aki(f["haplotag"], c["haplotype"], c["genotype"], fe["p"], fe["prob"]))
    return logp


def cross_optimize(cands, frags, with_genotype, counters):
    hg_inc = ht_inc = True
    iters = 0
    counters["calls"] += 1
    while hg_inc or ht_inc:
        counters["iters"] += 1
        tmp_tag = {}
        logp = pre_logp = 0.0
        for k, f in enumerate(frags):
            if not f["for_phasing"] or f.get("ds_skip") or f["haplotag"] == 0:
                continue
            pl = [fe for fe in f["list"] if fe["phase_site"]]
            if not pl:
                continue
            delta = [cands[fe["snp"]]["haplotype"] for fe in pl]
            eta = [cands[fe["snp"]]["genotype"] for fe in pl]
            ps = [fe["p"] for fe in pl]
            probs = [fe["prob"] for fe in pl]
            q = cal_sigma_delta_eta_log(f["haplotag"], delta, eta, ps, probs)
            qn = cal_sigma_delta_eta_log(-f["haplotag"], delta, eta, ps, probs)
            tmp_tag[k] = -f["haplotag"] if q < qn else f["haplotag"]
            # check_new_haplotag (phase.rs:278-314)
            logp += qn if q < qn else q
            pre_logp += q
        assert not (logp < pre_logp)
        for k, h in tmp_tag.items():
            frags[k]["haplotag"] = h
        if logp > pre_logp:
            ht_inc = hg_inc = True
        else:
            ht_inc = False
        tmp_hg = {}
        logp = pre_logp = 0.0
        for i, c in enumerate(cands):
            if not c["for_phasing"]:
                continue
            sigma, ps, probs = [], [], []
            for k in c["cover"]:
                f = frags[k]
                if not f["for_phasing"] or f.get("ds_skip") or f["haplotag"] == 0:
                    continue
                for fe in f["list"]:
                    if fe["snp"] == i and fe["phase_site"]:
                        ps.append(fe["p"])
                        probs.append(fe["prob"])
                        sigma.append(f["haplotag"])
            if not sigma:
                continue
            d, e = c["haplotype"], c["genotype"]
            q1 = cal_delta_eta_sigma_log(d, 0, sigma, ps, probs)
            q2 = cal_delta_eta_sigma_log(-d, 0, sigma, ps, probs)
            q3 = cal_delta_eta_sigma_log(d, 1, sigma, ps, probs)
            q4 = cal_delta_eta_sigma_log(d, -1, sigma, ps, probs)
            if with_genotype:
                mq = max(q1, q2, q3, q4)
                new = (d, 0) if q1 == mq else ((-d, 0) if q2 == mq else ((d, 1) if q3 == mq else (d, -1)))
            elif e == 0:
                new = (d, 0) if q1 == max(q1, q2) else (-d, 0)
            else:
                new = (d, 1) if q3 == max(q3, q4) else (d, -1)
            tmp_hg[i] = new
            # check_new_haplotype_genotype (phase.rs:316-355)
            logp += cal_delta_eta_sigma_log(new[0], new[1], sigma, ps, probs)
            pre_logp += cal_delta_eta_sigma_log(d, e, sigma, ps, probs)
        assert not (logp < pre_logp)
        for i, (d, e) in tmp_hg.items():
            cands[i]["haplotype"], cands[i]["genotype"] = d, e
        if logp > pre_logp:
            hg_inc = ht_inc = True
        else:
            hg_inc = False
        iters += 1
        if iters > 20:
            break
    return overall_probability(cands, frags)


def phase_enum(P, region, cands, frags, reads_in_region_index):
    """phase.rs:1097-1122 with the contract's random source for init_assignment (stream 1, call = configuration, idx = read index in the region)"""
    n = len(cands)
    assert n <= P["max_enum_snps"]
    rkey = region_key(region["tid"], region["start"])
    best, largest = None, -math.inf
    counters = dict(calls=0, iters=0)
    for cfg in range(1 << n):
        for i, c in enumerate(cands):
            c["haplotype"] = -1 if (cfg >> i) & 1 else 1
        for f in frags:
            if f["for_phasing"]:
                f["haplotag"] = -1 if uniform(P["seed"], rkey, RNG_INIT_SIGMA, cfg, reads_in_region_index[f["read"]]) < 0.5 else 1
        for c in cands:  # init_genotype, phase.rs:682-691
            c["genotype"] = {0: 1, 1: 0, 2: -1, 3: -1}.get(c["variant_type"], c["genotype"])
        prob = cross_optimize(cands, frags, True, counters)
        if prob > largest:
            largest = prob
            best = ([c["haplotype"] for c in cands], [c["genotype"] for c in cands], [f["haplotag"] for f in frags])
    for c, h, g in zip(cands, best[0], best[1]):
        c["haplotype"], c["genotype"] = h, g
    for f, t in zip(frags, best[2]):
        f["haplotag"] = t
    return counters


# ---------------------------------------------------------------- X3 --downsample   phase.rs:693-701, thread.rs:144-151
def _chacha_block(key, counter, rounds):
    """One ChaCha block (D. J. Bernstein), 64-bit block counter in words 12-13, zero stream id: what rand_chacha 0.3 generates."""
    M = 0xFFFFFFFF
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key) + [counter & M, (counter >> 32) & M, 0, 0]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M
        x[d] ^= x[a]
        x[d] = ((x[d] << 16) | (x[d] >> 16)) & M
        x[c] = (x[c] + x[d]) & M
        x[b] ^= x[c]
        x[b] = ((x[b] << 12) | (x[b] >> 20)) & M
        x[a] = (x[a] + x[b]) & M
        x[d] ^= x[a]
        x[d] = ((x[d] << 8) | (x[d] >> 24)) & M
        x[c] = (x[c] + x[d]) & M
        x[b] ^= x[c]
        x[b] = ((x[b] << 7) | (x[b] >> 25)) & M

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12), qr(1, 5, 9, 13), qr(2, 6, 10, 14), qr(3, 7, 11, 15)
        qr(0, 5, 10, 15), qr(1, 6, 11, 12), qr(2, 7, 8, 13), qr(3, 4, 9, 14)
    return [(a + b) & M for a, b in zip(x, st)]


class StdRng:
    """rand 0.8.5 StdRng (ChaCha12) seeded with SeedableRng::seed_from_u64 (rand_core 0.6: a PCG32 stream fills the key)."""

    def __init__(self, seed):
        state, key = seed & M64, []
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) & M64
            xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            key.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF)
        self.key, self.counter, self.buf = key, 0, []

    def next_u32(self):
        if not self.buf:
            self.buf = _chacha_block(self.key, self.counter, 12)
            self.counter += 1
        return self.buf.pop(0)

    def below(self, n):
        """UniformInt<u32>::sample_single(0, n): widening multiply, rejection zone (n << leading_zeros(n)) - 1"""
        zone = ((n << (32 - n.bit_length())) - 1) & 0xFFFFFFFF
        while True:
            m = self.next_u32() * n
            if (m & 0xFFFFFFFF) <= zone:
                return m >> 32


def downsample_fragments(n_fragments, depth, seed=2025):
    """Indices of the fragments marked `downsampled`: the first `depth` of SliceRandom::shuffle over 0..n (phase.rs:693-701)."""
    rng = StdRng(seed)
    idx = list(range(n_fragments))
    for i in range(n_fragments - 1, 0, -1):
        j = rng.below(i + 1)
        idx[i], idx[j] = idx[j], idx[i]
    return idx[:depth]


# ---------------------------------------------------------------- X2 imported candidates   candidate.rs:530-613
def import_external_candidates(region, fv, records, min_variant_qual=0.0):
    """records: {0-based position: (genotype class, quality)} of the region's contig (vcf.rs:400-462).  Returns the candidate list in
    the form candidates() returns it (the flags import never sets are False)."""
    cands = []
    position = region["start"] - 1
    for bf in fv:
        pos = position
        position += 1
        if pos not in records:
            continue
        gt, quality = records[pos]
        a1, c1, a2, c2 = two_major(bf)
        if f32(quality) < f32(min_variant_qual):
            continue
        total = bf["a"] + bf["c"] + bf["g"] + bf["t"]
        with_nan = lambda c: f32(f32(c) / f32(total)) if total else float("nan")  # noqa: E731  (0 / 0 in f32)
        c = dict(pos=pos, reference=bf["ref_base"], alleles=(a1, a2), freqs=(with_nan(c1), with_nan(c2)), depth=total,
                 variant_quality=float(f32(quality)), genotype_quality=float(f32(quality)), gp=[0.0, 0.0, 0.0], haplotype=0, phase_score=0.0,
                 rna_editing=False, dense=False, het_var=False, for_phasing=False, hom_var=False, cand_somatic=False, single=False, non_selected=False,
                 cover=[])
        if gt == 1:
            c.update(variant_type=1, genotype=0, for_phasing=True, het_var=True)
        elif gt == 2:
            c.update(variant_type=2, genotype=-1, for_phasing=True, hom_var=True)
        elif gt == 3:
            c.update(variant_type=3, genotype=-1, hom_var=True)
        else:  # 0/0 is built but never pushed; anything else is reported and skipped
            continue
        cands.append(c)
    return cands


# ---------------------------------------------------------------- R0 isolated regions   util.rs:236-332
def find_isolated_regions(ref_len, reads, min_mapq, min_read_length, divergence, truncation=False, truncation_coverage=200000):
    """One contig.  reads: iterable of dicts {mapq, l_seq, flag, de (float or None), pos, end} with [pos, end) the
    reference span of the alignment (record.reference_start() / reference_end()).  Returns [(start, end, max_coverage)],
    start 1-based inclusive, end 1-based exclusive, exactly as the per-position loop of util.rs:282-330 produces them,
    including its two quirks: a covered run of one position is not pushed and not forgotten (it becomes the start of the
    region that the next run closes), and max_coverage is taken before the push test."""
    depth = [0] * ref_len
    for r in reads:
        if r["mapq"] < min_mapq or r["l_seq"] < min_read_length or (r["flag"] & 0x4) or (r["flag"] & 0x100) or (r["flag"] & 0x800):  # util.rs:262-270
            continue
        if r["de"] is not None and f32(r["de"]) >= f32(divergence):  # util.rs:272-279
            continue
        for i in range(max(r["pos"], 0), min(r["end"], ref_len)):  # util.rs:281-285
            depth[i] += 1
    out = []
    region_start = region_end = -1
    max_coverage = 0
    for i in range(ref_len):  # util.rs:290-318
        if depth[i] > max_coverage:
            max_coverage = depth[i]
        if depth[i] == 0 or (truncation and depth[i] > truncation_coverage):
            if region_end > region_start:
                out.append((region_start + 1, region_end + 2, max_coverage))
                region_start = region_end = -1
                max_coverage = 0
        else:
            if region_start == -1:
                region_start = region_end = i
            else:
                region_end = i
    if region_end > region_start:  # util.rs:319-329
        out.append((region_start + 1, region_end + 2, max_coverage))
    return out


# ---------------------------------------------------------------- B0 phased BAM   thread.rs:307-361
def phased_bam(records, regions, haplotag_queue, phaseset_queue):
    """records: BAM-ordered dicts {qname, tid, pos, end, flag, tags: {tag: value}} (end = htslib bam_endpos);
    regions: [(tid, start, end)] in output order; the two queues are the (qname, value) pairs in the order the region
    workers pushed them.  Returns [(record index, HP or None, PS or None)] in the order the writer emits them; None
    means no aux field is pushed (or rust-htslib refused it because the tag was already there and the error was ignored)."""
    read_assignments = {}
    for q, v in haplotag_queue:  # thread.rs:308-316
        if q not in read_assignments:
            read_assignments[q] = v
    read_phasesets = {}
    for q, v in phaseset_queue:  # thread.rs:317-325
        if q not in read_phasesets:
            read_phasesets[q] = v
    out = []
    for tid, start, end in regions:  # thread.rs:330-358
        for i, r in enumerate(records):
            if r["tid"] != tid or not (r["pos"] < end and r["end"] > start):  # fetch((chr, start, end))
                continue
            if r["flag"] & 0x4 or r["flag"] & 0x100 or r["flag"] & 0x800:
                continue
            if r["pos"] + 1 < start or r["end"] + 1 > end:
                continue
            hp = ps = None
            if r["qname"] in read_assignments and read_assignments[r["qname"]] != 0 and "HP" not in r["tags"]:
                hp = read_assignments[r["qname"]]
            if r["qname"] in read_phasesets and "PS" not in r["tags"]:
                ps = read_phasesets[r["qname"]]
            out.append((i, hp, ps))
    return out
