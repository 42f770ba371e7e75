"""proof.py — Proof, Queries, FriProof, LayerProof, MerkleView and their wire format.

Mirrors src/proof.rs:43-66, src/trace.rs:37-66, src/fri.rs:71-125, src/merkle.rs:71-80.  The byte layout is
ark-serialize's (compressed mode): struct fields in declaration order; integers little-endian (usize as u64);
Vec<T> = u64 length + items; Option<T> = one tag byte (0 / 1) + item; a digest = the 32-byte slice, i.e. u64
length 32 + bytes (SerdeOutput delegates to the byte slice, src/utils.rs:552-560); an Fp = 8-byte LE canonical
integer; an Fq3 = c0 ‖ c1 ‖ c2.  (ark-serialize itself is not under /root/reference: SURVEY.md §8c lists these
conventions as restated from upstream.)

Field elements inside a Proof are canonical integers (Fp) or 3-tuples (Fq3).
"""
from dataclasses import dataclass, field
from typing import List, Optional

from .air import ProofOptions
from .channel import serialize_element


def _u64(v):
    return int(v).to_bytes(8, "little")


def _vec(items, enc):
    return _u64(len(items)) + b"".join(enc(x) for x in items)


def _digest(d):
    return _u64(32) + bytes(d)


@dataclass
class MerkleView:
    nodes: List[bytes]
    initial_leaves: List[bytes]
    sibling_leaves: List[bytes]
    height: int

    def to_bytes(self):
        return (_vec(self.nodes, _digest) + _vec(self.initial_leaves, _digest) + _vec(self.sibling_leaves, _digest)
                + int(self.height).to_bytes(4, "little"))


@dataclass
class LayerProof:
    flattenend_rows: list            # (sic) src/fri.rs:101
    merkle_proof: MerkleView
    commitment: bytes

    def to_bytes(self):
        return _vec(self.flattenend_rows, serialize_element) + self.merkle_proof.to_bytes() + _digest(self.commitment)


@dataclass
class FriProof:
    layers: List[LayerProof]
    remainder_coeffs: list

    def to_bytes(self):
        return _vec(self.layers, LayerProof.to_bytes) + _vec(self.remainder_coeffs, serialize_element)


@dataclass
class Queries:
    base_trace_values: list
    extension_trace_values: list
    composition_trace_values: list
    base_trace_proof: MerkleView
    extension_trace_proof: Optional[MerkleView]
    composition_trace_proof: MerkleView

    def to_bytes(self):
        ext = b"\x00" if self.extension_trace_proof is None else b"\x01" + self.extension_trace_proof.to_bytes()
        return (_vec(self.base_trace_values, serialize_element) + _vec(self.extension_trace_values, serialize_element)
                + _vec(self.composition_trace_values, serialize_element) + self.base_trace_proof.to_bytes() + ext
                + self.composition_trace_proof.to_bytes())


@dataclass
class Proof:
    options: ProofOptions
    trace_len: int
    base_trace_commitment: bytes
    extension_trace_commitment: Optional[bytes]
    composition_trace_commitment: bytes
    fri_proof: FriProof
    pow_nonce: int
    trace_queries: Queries
    execution_trace_ood_evals: list
    composition_trace_ood_evals: list
    timings: dict = field(default_factory=dict, compare=False)      # per-phase wall clock, like the reference's println!s

    def to_bytes(self):
        ext = (b"\x00" if self.extension_trace_commitment is None
               else b"\x01" + _digest(self.extension_trace_commitment))
        return (self.options.to_bytes() + _u64(self.trace_len) + _digest(self.base_trace_commitment) + ext
                + _digest(self.composition_trace_commitment) + self.fri_proof.to_bytes() + _u64(self.pow_nonce)
                + self.trace_queries.to_bytes() + _vec(self.execution_trace_ood_evals, serialize_element)
                + _vec(self.composition_trace_ood_evals, serialize_element))
