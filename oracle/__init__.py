"""CPU oracle for the WAE-training / CLaSS-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this
directory; it may be used by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`, and only as the checker
or the timed CPU baseline -- never as a fallback for the CUDA path.

The oracle is a restatement, in torch-CPU / numpy, of the arithmetic the
reference (IBM/controlled-peptide-generation @ 1ba3ce8) performs on its hot
path.  Every function cites the reference file:line it follows.  The reference
has no golden vectors of its own ("parity unpinned" by the reference's tests),
so the pins are generated from the *live* reference, imported unmodified from
/root/reference in the build container by `oracle/gen_golden.py`, and committed
under `tests/golden/`.  `tests/test_oracle_golden.py` checks this restatement
against those vectors on every run (no GPU needed).

Third-party arithmetic the reference delegates to (not under /root/reference):
torch (pinned 1.7.1 by amp_gen.yml:8; 2.11.0 here), scikit-learn (unpinned;
1.9.0 here), numpy RNG (unpinned; 2.3.5 here).  Goldens record these versions.
"""
