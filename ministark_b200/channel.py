"""channel.py — Fiat–Shamir plumbing of the prover: PublicCoinImpl<F, Sha256HashFn> (src/random.rs:91-196)
and ProverChannel (src/channel.rs:17-134).

Host-side and tiny (a few hundred SHA-256 calls per proof): hashlib does the hashing.  The one heavy item,
proof-of-work grinding, goes to the GPU (Context.pow_grind -> ms_pow_grind_sha256) with the deterministic
smallest-nonce rule.

Field elements are canonical integers (Fp) or 3-tuples (Fq3).  Conventions that live in crates outside
/root/reference (ark-ff-optimized `Fp::rand`, rand 0.8.5 `gen_range`) are restated from SURVEY.md §8c and
flagged there as unverifiable in this environment; everything the reference itself defines (byte order of the
coin, seed/counter hashing, reseeding) follows src/random.rs line by line.
"""
import hashlib

P = 2**64 - 2**32 + 1
_RINV = pow(2**64, -1, P)


def _sha(*chunks):
    h = hashlib.sha256()
    for c in chunks:
        h.update(c)
    return h.digest()


def serialize_element(v):
    """Field::serialize_uncompressed: 8-byte LE canonical integer per base-field limb (c0, c1, c2 for Fq3)"""
    if isinstance(v, (tuple, list)):
        return b"".join(int(c).to_bytes(8, "little") for c in v)
    return int(v).to_bytes(8, "little")


def hash_elements(elements):
    """Sha256HashFn::hash_elements (src/hash.rs:92-99)"""
    return _sha(b"".join(serialize_element(e) for e in elements))


def merge(a, b):
    return _sha(a, b)                                           # src/hash.rs:77-82


def merge_with_int(seed, value):
    return _sha(seed, int(value).to_bytes(8, "big"))            # src/hash.rs:84-89


def leading_zeros(digest):
    z = 0
    for byte in digest:                                         # src/random.rs:183-195
        lz = 8 - byte.bit_length()
        z += lz
        if lz != 8:
            break
    return z


class PublicCoin:
    """PublicCoinImpl (src/random.rs:91-181).  `ext=False`: Field = Fp; `ext=True`: Field = Fq3."""

    def __init__(self, seed, ext=False):
        self.seed, self.counter, self.bytes, self.ext = bytes(seed), 0, b"", ext

    def _reset(self, seed):
        self.seed, self.counter, self.bytes = seed, 0, b""

    def reseed_with_digest(self, d):
        self._reset(merge(self.seed, d))

    def reseed_with_field_elements(self, vals):
        for v in vals:                                          # one hash_elements + merge per element (:104-109)
            self._reset(merge(self.seed, hash_elements([v])))

    reseed_with_field_element_vector = reseed_with_field_elements

    def reseed_with_int(self, v):
        self._reset(merge_with_int(self.seed, v))

    def verify_proof_of_work(self, bits, nonce):
        return leading_zeros(merge_with_int(self.seed, nonce)) >= bits

    # ---- the RngCore view (src/random.rs:150-181): bytes are popped from the END of hash(seed || counter)
    def _next_byte(self):
        if not self.bytes:
            self.counter += 1
            self.bytes = merge_with_int(self.seed, self.counter)
        b = self.bytes[-1]
        self.bytes = self.bytes[:-1]
        return b

    def next_u64(self):
        if not self.bytes:
            self.counter += 1
            self.bytes = merge_with_int(self.seed, self.counter)
        if len(self.bytes) >= 8:        # eight pops from the end, first pop = most significant byte: the tail read little-endian
            v = int.from_bytes(self.bytes[-8:], "little")
            self.bytes = self.bytes[:-8]
            return v
        return int.from_bytes(bytes(self._next_byte() for _ in range(8)), "big")

    def _draw_fp(self):
        # ark-ff-optimized fp64 `Standard::sample`: take next_u64 until it is < p and use it AS THE MONTGOMERY WORD
        # ("known only from upstream memory"); returned here as the canonical integer it represents.
        while True:
            w = self.next_u64()
            if w < P:
                return w * _RINV % P

    def draw(self):
        if self.ext:
            return (self._draw_fp(), self._draw_fp(), self._draw_fp())
        return self._draw_fp()

    def draw_queries(self, max_n, domain_size):
        """BTreeSet of max_n draws of gen_range(0..domain_size) (src/random.rs:139-141) — rand 0.8.5
        UniformInt::sample_single_inclusive for a 64-bit usize: widening multiply with zone rejection."""
        rng_range = domain_size                                 # high - low + 1 with inclusive high = domain_size - 1
        lz = 64 - rng_range.bit_length()
        zone = ((rng_range << lz) - 1) & (2**64 - 1)
        out = set()
        for _ in range(max_n):
            while True:
                m = self.next_u64() * rng_range
                if (m & (2**64 - 1)) <= zone:
                    out.add(m >> 64)
                    break
        return sorted(out)


class ProverChannel:
    """src/channel.rs:17-134"""

    def __init__(self, air, public_coin, ctx):
        self.air, self.public_coin, self.ctx = air, public_coin, ctx
        self.base_trace_commitment = None
        self.extension_trace_commitment = None
        self.composition_trace_commitment = None
        self.fri_layer_commitments = []
        self.fri_remainder_coeffs = []
        self.execution_trace_ood_evals = []
        self.composition_trace_ood_evals = []
        self.pow_nonce = 0

    def commit_base_trace(self, root):
        self.public_coin.reseed_with_digest(root)
        self.base_trace_commitment = root

    def commit_extension_trace(self, root):
        self.public_coin.reseed_with_digest(root)
        self.extension_trace_commitment = root

    def commit_composition_trace(self, root):
        self.public_coin.reseed_with_digest(root)
        self.composition_trace_commitment = root

    def get_ood_point(self):
        return self.public_coin.draw()

    def send_ood_evals(self, execution_trace_oods, composition_trace_oods):
        self.public_coin.reseed_with_field_elements(list(execution_trace_oods) + list(composition_trace_oods))
        self.execution_trace_ood_evals = list(execution_trace_oods)
        self.composition_trace_ood_evals = list(composition_trace_oods)

    def grind_fri_commitments(self):
        bits = self.air.options.grinding_factor
        if bits == 0:
            return
        nonce = self.ctx.pow_grind(self.public_coin.seed, bits)      # GPU, smallest nonce >= 1
        assert self.public_coin.verify_proof_of_work(bits, nonce)
        self.pow_nonce = nonce
        self.public_coin.reseed_with_int(nonce)

    def get_fri_query_positions(self):
        n = self.air.trace_len * self.air.lde_blowup_factor()
        return self.public_coin.draw_queries(self.air.options.num_queries, n)

    # fri::ProverChannel (src/channel.rs:122-140)
    def commit_fri_layer(self, root):
        self.public_coin.reseed_with_digest(root)
        self.fri_layer_commitments.append(root)

    def commit_remainder(self, coeffs):
        self.public_coin.reseed_with_field_element_vector(coeffs)
        self.fri_remainder_coeffs = list(coeffs)

    def draw_fri_alpha(self):
        return self.public_coin.draw()
