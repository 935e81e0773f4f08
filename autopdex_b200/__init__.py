"""autopdex_b200 -- B200 (sm_100a) backend for AutoPDEx's hot path:
sparse residual/tangent assembly followed by the Newton linear solve.

Host side mirrors the reference interface for that path (`solver.solver`,
`solver.adaptive_load_stepping`, `assembler.assemble_*`, the `models` / `spaces` / `seeder` /
`mesher` names the built-in models need); all arithmetic runs in libapdx_b200.so.
"""
__version__ = "0.1.0"
