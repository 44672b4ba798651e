"""CPU oracle for the PfoTGNRec hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy / CPU-torch restatement of the reference's
algorithm for the path named in BASELINE.json.  It is
the checker the CUDA path is compared against.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; nothing under `pfotgnrec_b200/` does, and the product
path raises if its CUDA library is missing instead of falling back to this.

Pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the unmodified reference modules
imported from /root/reference in the build container; the generating script is
`oracle/make_golden.py` and the vectors live in `tests/golden/`.
Documented deviations from the reference:
  (i)  random draws come from a counter-based Philox4x32-10 stream
       (`oracle/philox.py`) instead of numpy's global MT19937;
  (ii) the three `argsort` call sites use a stable sort;
  (iii) attention dropout is 0 (or eval mode) in every parity case.
"""
