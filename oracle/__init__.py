"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Mixer-CLIP training hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the timed CPU baseline.
The product path (``clip_mixer_b200``) never imports this package and raises when the
CUDA library is missing.

Parity pinning: the reference's own tests hold no golden vector for the Mixer path
(SURVEY.md section 4), so the restatement in ``mixer_clip_oracle.py`` is pinned against
outputs of the reference module itself, generated in the build container by
``oracle/make_golden.py`` (which imports ``/root/reference/training/clip/model.py`` by
file path) and committed under ``tests/golden/``.
"""
