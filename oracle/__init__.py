"""CPU oracle for the log-mel + Whisper-encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package
(`taiwan-tongues-asr-ce_b200/`) may import this package.  The only callers are
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py`, and there only as the checker / the CPU arm.

Parity pin: the reference repository holds no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against the installed Hugging Face
`transformers` implementation that the reference calls
(`train_asr.py:518-527,607-616`): `tests/golden/*.npz` were produced by
`oracle/gen_golden.py` from that implementation in the build container, and
`tests/test_oracle_*.py` additionally re-check the oracle live against
`transformers` whenever it is importable.
"""
