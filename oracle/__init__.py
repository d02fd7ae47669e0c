"""CPU oracle for the supervised-autoencoder + MLP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product
(``ae_b200``) never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference itself: ``oracle/make_golden.py`` executes the
reference's own notebook cells (NB:499-525, NB:607-635, NB:685-702, NB:2970-2987 and the
literal train-step lines NB:2652-2654 / NB:2676-2684) in this container and commits their
outputs under ``tests/golden/``; ``tests/test_oracle_pin.py`` checks this restatement against
those vectors.  The input transforms (NB:361-368, NB:386-395) are restated in
``oracle/augment_port.py`` and pinned the same way (``oracle/make_golden_augment.py`` runs the
reference's own torchvision ``Compose`` objects).
"""
