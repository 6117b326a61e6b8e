"""CPU oracle for the NAF hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
(``neuralvolumetricreconstructionformedicalimages_b200``) never does.

Parity status: PINNED.
  * ``nafb_oracle.c`` (hash-grid op) is bit-identical to the reference's own kernel
    text compiled for the host (``oracle/_ref``, built by ``oracle/build_ref.sh``
    from ``/root/reference/src/encoder/hashencoder/src/hashencoder.cu:30-298``) --
    checked by ``tests/test_oracle_pin.py`` when ``oracle/_ref`` exists, and against
    the committed fixtures under ``tests/golden/`` everywhere.
  * ``naf.py`` (sampling, MLP, ray integral, loss, geometry) is checked against
    outputs of the reference's Python modules imported from ``/root/reference`` in
    the build container; the generating script is ``tests/golden/generate_golden.py``
    and its outputs are the ``tests/golden/*.npz`` fixtures.
"""
from .hashgrid import (  # noqa: F401
    OracleHashEncoder,
    build_oracle,
    have_ref,
    level_offsets,
    oracle_corners,
    oracle_hash_backward,
    oracle_hash_forward,
    ref_hash_backward,
    ref_hash_forward,
)
