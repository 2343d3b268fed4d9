"""Synthetic inputs shared by tests and bench.py (numpy only)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def unpack_2bit(words, n):
    """uint64 words (base i at bits [2i,2i+2)) -> uint8 codes[n]."""
    words = np.asarray(words, dtype=np.uint64)
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))
    codes = ((words[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8).reshape(-1)
    return codes[:n]


def kmer_words_from_codes(codes, k):
    """all k-mer words of a code sequence (vectorised): word[p] = sum codes[p+i] << 2i."""
    n = len(codes) - k + 1
    if n <= 0:
        return np.zeros(0, dtype=np.uint64)
    w = np.zeros(n, dtype=np.uint64)
    c = codes.astype(np.uint64)
    for i in range(k):
        w |= c[i:i + n] << np.uint64(2 * i)
    return w


def sample_reads(ref_codes, n_reads, read_len, seed, frac_ref=0.5, sub_rate=0.0, n_rate=0.0, ragged=False):
    """Reads as ASCII bytes + offsets.  A fraction frac_ref is sampled from ref_codes at a uniform
    start on a random strand (with i.i.d. substitutions at sub_rate), the rest is uniform random ACGT.
    n_rate injects 'N' bytes; ragged draws lengths in [0, read_len]."""
    rng = np.random.default_rng(seed)
    if ragged:
        lens = rng.integers(0, read_len + 1, size=n_reads)
    else:
        lens = np.full(n_reads, read_len, dtype=np.int64)
    offs = np.zeros(n_reads + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    total = int(offs[-1])
    codes = rng.integers(0, 4, size=total, dtype=np.uint8)
    is_ref = rng.random(n_reads) < frac_ref
    L = len(ref_codes)
    for r in np.nonzero(is_ref)[0]:
        ln = int(lens[r])
        if ln == 0 or ln > L:
            continue
        s = int(rng.integers(0, L - ln + 1))
        seg = ref_codes[s:s + ln]
        if rng.integers(0, 2):
            seg = COMP[seg[::-1]]
        codes[int(offs[r]):int(offs[r]) + ln] = seg
    if sub_rate > 0:
        m = rng.random(total) < sub_rate
        codes[m] = (codes[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
    bases = ACGT[codes].copy()
    if n_rate > 0:
        m = rng.random(total) < n_rate
        bases[m] = ord("N")
    # sprinkle lower case: the path is case-insensitive (src/pf1/dense_index.rs:180-183)
    m = rng.random(total) < 0.1
    bases[m] |= 0x20
    return bases, offs


def sample_reads_fast(ref_codes, n_reads, read_len, seed, frac_ref=0.5, sub_rate=0.0):
    """Vectorised uniform-length generator for large batches (bench.py): same mix as sample_reads."""
    rng = np.random.default_rng(seed)
    L = len(ref_codes)
    is_ref = rng.random(n_reads) < frac_ref
    starts = rng.integers(0, L - read_len + 1, size=n_reads)
    strand = rng.integers(0, 2, size=n_reads).astype(bool)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    codes = ref_codes[idx]
    rc = COMP[codes[:, ::-1]]
    codes = np.where(strand[:, None], rc, codes)
    rnd = rng.integers(0, 4, size=(n_reads, read_len), dtype=np.uint8)
    codes = np.where(is_ref[:, None], codes, rnd)
    if sub_rate > 0:
        m = rng.random((n_reads, read_len)) < sub_rate
        codes = np.where(m, (codes + rng.integers(1, 4, size=(n_reads, read_len), dtype=np.uint8)) & 3, codes)
    return ACGT[codes.astype(np.uint8)].reshape(-1).copy()


def synthetic_unitigs(n_unitigs, mean_extra, k, seed):
    """uniform-random unitig set: lengths = k + Geometric(mean mean_extra); returns (codes, accum)."""
    rng = np.random.default_rng(seed)
    lens = k + rng.geometric(1.0 / max(mean_extra, 1), size=n_unitigs) - 1
    accum = np.zeros(n_unitigs + 1, dtype=np.uint64)
    accum[1:] = np.cumsum(lens)
    codes = rng.integers(0, 4, size=int(accum[-1]), dtype=np.uint8)
    return codes, accum


# ------------------------------------------------------------------------------------------------
# large synthetic inputs (bench.py config 5): packed random unitigs on the host, reads on the device
# ------------------------------------------------------------------------------------------------
def synthetic_unitigs_packed(n_unitigs, mean_extra, k, seed):
    """Uniform-random unitig set without materialising per-base codes: random u64 words ARE random
    2-bit sequences.  lengths = k + Geometric(mean mean_extra).  Returns (words, n_bases, accum)."""
    rng = np.random.default_rng(seed)
    lens = (k + rng.geometric(1.0 / max(mean_extra, 1), size=n_unitigs) - 1).astype(np.uint64)
    accum = np.zeros(n_unitigs + 1, dtype=np.uint64)
    np.cumsum(lens, out=accum[1:])
    n_bases = int(accum[-1])
    nw = (2 * n_bases + 63) // 64
    words = rng.integers(0, np.iinfo(np.uint64).max, size=nw, dtype=np.uint64, endpoint=True)
    if (2 * n_bases) & 63:
        words[-1] &= np.uint64((1 << ((2 * n_bases) & 63)) - 1)
    return words, n_bases, accum


def device_reads_from_packed(torch, useq_words_dev, n_bases, n_reads, read_len, gen, frac_ref, sub_rate, chunk=1_000_000):
    """ASCII reads on the device: frac_ref sampled from the packed sequence at a uniform start on a
    random strand (i.i.d. substitutions at sub_rate), the rest uniform random.  useq_words_dev: int64 view."""
    dev = useq_words_dev.device
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    bases = torch.empty(n_reads * read_len, dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for r0 in range(0, n_reads, chunk):
        n = min(chunk, n_reads - r0)
        starts = torch.randint(0, n_bases - read_len + 1, (n,), generator=gen, device=dev)
        p = starts[:, None] + ar[None, :]
        codes = ((useq_words_dev[p >> 5] >> ((p & 31) << 1)) & 3).to(torch.uint8)
        strand = torch.rand(n, generator=gen, device=dev) < 0.5
        rc = (3 - codes.flip(1))
        codes = torch.where(strand[:, None], rc, codes)
        is_ref = torch.rand(n, generator=gen, device=dev) < frac_ref
        rnd = torch.randint(0, 4, (n, read_len), generator=gen, device=dev, dtype=torch.uint8)
        codes = torch.where(is_ref[:, None], codes, rnd)
        if sub_rate > 0:
            m = torch.rand((n, read_len), generator=gen, device=dev) < sub_rate
            codes = torch.where(m, (codes + torch.randint(1, 4, (n, read_len), generator=gen, device=dev, dtype=torch.uint8)) & 3, codes)
        bases[r0 * read_len:(r0 + n) * read_len] = acgt[codes.long()].reshape(-1)
    return bases


def device_kmers_from_packed(torch, useq_words_dev, n_bases, n, k, gen, frac_pos):
    """forward k-mer words on the device: frac_pos sampled from the packed sequence (random strand is
    left to the index: a window IS a k-mer of the set unless it straddles a unitig boundary), rest random."""
    dev = useq_words_dev.device
    pos = torch.randint(0, n_bases - k + 1, (n,), generator=gen, device=dev)
    wi, sh = pos >> 5, (pos & 31) << 1
    lo = useq_words_dev[wi]
    hi = useq_words_dev[wi + 1]
    # logical funnel shift on int64 storage: mask off the sign-extension of the arithmetic >>
    keep = torch.where(sh == 0, torch.full_like(sh, -1), (torch.ones_like(sh) << (64 - sh).clamp(max=63)) - 1)
    lo_s = (lo >> sh) & keep
    hi_s = torch.where(sh == 0, torch.zeros_like(hi), hi << (64 - sh))
    w = lo_s | hi_s
    mask = (1 << (2 * k)) - 1 if k < 32 else -1
    w = w & mask
    rnd = torch.randint(0, 1 << 62, (n,), generator=gen, device=dev) & mask
    is_pos = torch.rand(n, generator=gen, device=dev) < frac_pos
    return torch.where(is_pos, w, rnd)


def reads_from_packed(words, n_bases, n_reads, read_len, seed, frac_ref, sub_rate, codes=None):
    """numpy twin of device_reads_from_packed (the CPU arms of bench.py): the same mix drawn with numpy's generator.
    `codes` = unpack_2bit(words, n_bases) if the caller already holds it (one byte per base)."""
    rng = np.random.default_rng(seed)
    if codes is None:
        codes = unpack_2bit(words, n_bases)
    rows = np.lib.stride_tricks.sliding_window_view(codes, read_len)
    out = rng.integers(0, 4, size=(n_reads, read_len), dtype=np.uint8)
    is_ref = np.nonzero(rng.random(n_reads) < frac_ref)[0]
    starts = rng.integers(0, n_bases - read_len + 1, size=len(is_ref))
    ref = rows[starts]
    strand = rng.random(len(is_ref)) < 0.5
    ref[strand] = COMP[ref[strand][:, ::-1]]
    out[is_ref] = ref
    out = out.reshape(-1)
    if sub_rate > 0:
        n_sub = rng.binomial(out.size, sub_rate)
        at = rng.integers(0, out.size, size=n_sub)
        out[at] = (out[at] + rng.integers(1, 4, size=n_sub, dtype=np.uint8)) & 3
    return ACGT[out]
