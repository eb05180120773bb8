"""CPU oracle for the osu-diffusion DiT denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the timed CPU arm), never as the implementation behind the drop-in modules in
``osu-diffusion_b200/``.

The oracle is a functional restatement (plain torch on CPU, fp32 by default,
fp64 on request) of the reference algorithm; every function cites the
reference file:line it follows.  It is pinned against outputs of the unmodified
reference modules, generated in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/``
(the reference itself ships no tests or golden vectors: SURVEY.md F12).
"""
