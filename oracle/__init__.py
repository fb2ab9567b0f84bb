"""CPU oracle for the Voxtral-Mini-4B Q4_0 hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement of the reference's
arithmetic (TrevorS/voxtral-mini-realtime-rs @ ffad466).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import it -- and only as the checker, never as the product path.  The
product (``voxtral_mini_realtime_rs_b200``) never imports anything from here and
fails loudly when its CUDA library is missing.

Parity status: the reference cannot be built or imported in this environment (no
Rust toolchain, no Burn/CubeCL/wgpu, no model weights, no ``.npy`` fixtures -- see
SURVEY.md F1-F3).  The oracle is therefore pinned against every *closed-form* known
answer the reference's own tests contain (``tests/test_oracle_pins.py`` lists each
with its reference file:line).  For the pieces whose arithmetic lives in absent
third-party crates (Burn softmax/conv/matmul, rustfft) and whose reference tests need
absent ``.npy`` fixtures, the header of the respective module says "parity unpinned"
and DESIGN.md repeats it.
"""
