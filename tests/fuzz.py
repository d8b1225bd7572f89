"""Adversarial synthetic inputs for the parity tests: small random annotations (overlapping features, shared
boundaries, multi-element Order lines with upstream/downstream elements, stranded elements) and hit streams
that stress the per-read countdown (complete, truncated and over-long groups, inconsistent NH, names that
come back later, NH 0/1 records inside runs, degenerate intervals, unknown chromosomes)."""
import numpy as np

from mmannot_b200 import host


def make_elements(rng, n_elements=None, wide=False):
    """Random element table: Order lines of 1-3 elements, some stranded, some upstream/downstream."""
    if n_elements is None:
        n_elements = int(rng.integers(40, 60)) if wide else int(rng.integers(3, 20))
    line, strand, vic = [], [], []
    cur = 0
    while len(line) < n_elements:
        k = int(rng.choice([1, 1, 1, 2, 3]))
        k = min(k, n_elements - len(line))
        for _ in range(k):
            line.append(cur)
            strand.append(int(rng.choice([0, 0, 1, 2])))
            vic.append(int(rng.choice([0, 0, 0, 1, 2])) if k > 1 else int(rng.choice([0, 0, 0, 0, 1, 2])))
        cur += 1
    return host.ElementTable(np.array(line, np.uint16), np.array(strand, np.uint8), np.array(vic, np.uint8))


def make_features(rng, et, n_chr=3, n_feat=400, extent=20000, max_len=3000):
    """Random typed intervals sorted by (chromosome, start) with a stable sort; many share boundaries."""
    chr_ = rng.integers(0, n_chr, n_feat)
    grid = int(rng.choice([1, 1, 10, 50]))
    start = (rng.integers(1, extent, n_feat) // grid) * grid + 1
    length = np.where(rng.random(n_feat) < 0.7, rng.integers(1, 300, n_feat), rng.integers(1, max_len, n_feat))
    length = (length // grid) * grid + int(rng.choice([0, 1]))
    end = start + np.maximum(length, 0)
    type_ = rng.integers(0, et.n_elements, n_feat)
    strand = rng.integers(1, 3, n_feat)
    # leave one chromosome empty now and then
    if n_chr > 1 and rng.random() < 0.3:
        keep = chr_ != n_chr - 1
        chr_, start, end, type_, strand = chr_[keep], start[keep], end[keep], type_[keep], strand[keep]
    order = np.lexsort((start, chr_))  # lexsort is stable: ties keep the generation order
    return host.FeatureArrays(chr_[order], start[order], end[order], type_[order], strand[order], n_chr)


def _positions(rng, feats, n, extent, max_read):
    """Hit intervals: half of them hug feature boundaries."""
    chr_ = rng.integers(0, feats.n_chr, n).astype(np.uint32)
    start = rng.integers(1, extent + 2000, n)
    near = rng.random(n) < 0.6
    if feats.n:
        pick = rng.integers(0, feats.n, n)
        b = np.where(rng.random(n) < 0.5, feats.start[pick], feats.end[pick]).astype(np.int64)
        start = np.where(near, np.maximum(1, b + rng.integers(-40, 5, n)), start)
        chr_ = np.where(near, feats.chr[pick], chr_).astype(np.uint32)
    length = rng.integers(1, max_read + 1, n)
    end = start + length - 1
    # degenerate intervals (empty CIGAR: end = start - 1) and unknown chromosomes
    deg = rng.random(n) < 0.01
    end = np.where(deg, start - 1, end)
    unk = rng.random(n) < 0.01
    chr_ = np.where(unk, 0x00FFFFFF, chr_).astype(np.uint32)
    strand = (rng.random(n) < 0.5).astype(np.uint32) << np.uint32(31)
    return start.astype(np.uint32), end.astype(np.uint32), (chr_ | strand).astype(np.uint32)


def make_hits(rng, feats, n_reads=3000, extent=20000, max_nh=8, max_read=60, messy=0.15, key_pool=None):
    """Name-grouped stream of reads; `messy` = fraction of reads whose group is not the clean NH-records run."""
    nh_list, key_list = [], []
    next_key = 1
    pool = [] if key_pool is None else list(key_pool)
    for _ in range(n_reads):
        nh = int(rng.integers(1, max_nh + 1)) if rng.random() < 0.6 else 1
        if pool and rng.random() < 0.05:
            key = int(rng.choice(pool))  # a name that comes back later in the file
        else:
            key = next_key
            next_key += 1
            if rng.random() < 0.05:
                pool.append(key)
        n_rec = nh
        nhs = [nh] * n_rec
        if rng.random() < messy:
            kind = int(rng.integers(0, 5))
            if kind == 0:
                n_rec = int(rng.integers(1, nh + 1)); nhs = [nh] * n_rec            # truncated
            elif kind == 1:
                n_rec = nh + int(rng.integers(1, 4)); nhs = [nh] * n_rec            # too many records
            elif kind == 2:
                nhs = [int(rng.integers(0, max_nh + 1)) for _ in range(n_rec)]     # inconsistent NH (0 and 1 included)
            elif kind == 3:
                n_rec = nh + 1; nhs = [nh] * nh + [1]                                # a NH=1 record inside the run
            else:
                n_rec = 2 * nh; nhs = [nh] * n_rec                                   # two reads of the same name back to back
        nh_list.extend(nhs)
        key_list.extend([key] * len(nhs))
    n = len(nh_list)
    start, end, meta = _positions(rng, feats, n, extent, max_read)
    keys = np.array(key_list, np.uint64)
    # spread the keys over 64 bits; keep two special values in play
    keys = keys * np.uint64(0x9E3779B97F4A7C15)
    if n > 10:
        keys[keys == keys[n // 2]] = np.uint64(0xFFFFFFFFFFFFFFFF)
        keys[keys == keys[n // 3]] = np.uint64(0)
    return host.Hits(start, end, meta, np.array(nh_list, np.uint32), keys)


def shuffle_hits(rng, hits, block=1):
    """Coordinate-sorted-like disorder: the records of a read are no longer adjacent."""
    n = hits.n
    perm = rng.permutation(n) if block == 1 else np.concatenate([rng.permutation(np.arange(a, min(n, a + block))) for a in range(0, n, block)])
    return host.Hits(hits.start[perm], hits.end[perm], hits.meta[perm], hits.nh[perm], hits.read_key[perm])
