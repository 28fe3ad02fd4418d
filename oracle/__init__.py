"""CPU oracle for the SCAL-SDT LoRA train-step hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``scal_sdt_b200/`` imports it; the product path runs hand-written CUDA and
fails loudly when the CUDA library is missing.

Each function restates one piece of the reference (citations are relative to
``/root/reference``):

=====================  =====================================================
oracle module          follows
=====================  =====================================================
``lora_ref``           ``modules/lora.py:12-27`` + loralib==0.1 [ext]
``diffusion_ref``      ``modules/model.py:289-316,318-348`` + diffusers
                       ``DDIMScheduler.add_noise/get_velocity`` [ext]
``ema_ref``            ``modules/ema.py:9-140``
``reference_shim``     runs the reference's *own* ``modules/ema.py``,
                       ``modules/dataset/bucket.py`` and
                       ``modules/utils/torch/module.py`` (only in the build
                       container, to create ``tests/golden/``)
=====================  =====================================================

Pinning status
--------------
* EMA, bucket/rank sharding, target-selection DSL: PINNED -- golden files under
  ``tests/golden/`` were produced by executing the reference's own source
  (``oracle/make_golden.py``), and the restatements are checked against them.
* LoRA arithmetic (loralib 0.1) and DDIM ``add_noise``/``get_velocity``
  (unpinned diffusers fork): the arithmetic lives in third-party packages that
  are absent from ``/root/reference`` and from this image, and the reference
  ships no tests or golden vectors: PARITY UNPINNED by the reference.  The
  restatement follows the published algorithm of the pinned version and is
  cross-checked by closed-form identities (merged-weight equivalence, fp64
  autograd, ``alpha_bar`` analytic properties) in ``tests/test_oracle.py``.
"""
