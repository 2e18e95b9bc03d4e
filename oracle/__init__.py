"""CPU oracle for the generation-and-corruption hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``interpolated_diffusion_b200`` (the
product) may import this package.  The only legitimate users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- and there only as the checker / the CPU arm, never as the
thing that is shipped or measured as the GPU path.

Every function is a restatement (numpy for the integer / byte / elementwise
fp32 work, torch-CPU fp32 for the dense denoiser math) of the algorithm in
EquilibriaW/Interpolated_Diffusion and cites the ``file:line`` it follows.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the
live Python reference in the build container (``tests/golden/make_golden.py``
is the committed generator); ``tests/test_oracle_golden.py`` checks every
oracle function against them (bit-exact for masks / indices / interpolation /
DDIM arithmetic, 1e-5 for the dense model forwards).  One exception, stated in
its docstring: the batched chunk loop ``generate.generate_causal_chunked`` is
restated from the source (the live script needs the D4RL datasets); every
function it composes is pinned.
"""
