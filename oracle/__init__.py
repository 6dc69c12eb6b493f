"""CPU oracle for the neuroclear hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy / torch-CPU fp32 restatement of the reference's algorithm for the diced ``unet_deconv`` inference
path (geometry, dice, assemble, unet) and for the apollo training iteration (unet gradients, deeplinear,
discriminator, mip, apollo_step = the whole ``optimize_parameters()``), plus the training-data augmentation
(augment).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the checker or the timed CPU
baseline — never as a fallback for the CUDA path (``neuroclear_b200`` raises when its extension is missing).

Parity status: PINNED.  The reference ships no tests or golden vectors, so every function here
was checked against the unmodified reference modules imported from /root/reference in the build container
(``oracle/make_golden.py``); the inputs/outputs of those runs are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle_golden.py`` / ``tests/test_augment_host.py`` on every run (the training iteration
and the augmentation reproduce the reference's outputs bit for bit).  One piece is NOT pinned by reference output:
``skimage.exposure.rescale_intensity`` (scikit-image 0.18.3, not installed and not vendored) is restated from its
published algorithm in ``assemble.rescale_intensity`` — see the note there.
"""
