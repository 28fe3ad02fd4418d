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
* EMA, bucket / rank sharding, DreamBooth sampler, target-selection DSL: PINNED -- golden files under ``tests/golden/``
  were produced by executing the reference's own source (``oracle/make_golden.py``); restatements and product host
  code are checked against them.
* The reference's own GLUE around the third-party arithmetic: PINNED since round 2.  ``oracle/reference_shim.py``
  executes ``modules/lora.py`` (``get_lora``) and, by AST extraction, ``modules/model.py``'s ``get_optimizer``,
  ``config_module`` and ``LatentDiffusionModel`` (``_denoise_loss``, ``training_step``, ``on_save_checkpoint``) with
  stand-ins for ``loralib`` / ``diffusers`` / ``pytorch_lightning`` / ``omegaconf``.  Recorded in ``lora_glue.pt``,
  ``denoise_steps.pt``, ``config_module.json`` and checked by ``tests/test_reference_glue.py`` (CPU: oracle + product host
  code; GPU: the CUDA path): the isinstance switch, weight / bias aliasing, the int32 ``lora_alpha`` buffer, the state-dict
  keys, the order of the random draws, the ``"v"`` spelling of the target switch, the NaN guards and their messages, the
  ``chunk`` / ``mean`` prior-preservation reduction, the param groups, the LR / weight-decay scaling, the checkpoint keys.
* RESIDUE THAT STAYS UNPINNED BY THE REFERENCE: the arithmetic *inside* loralib-0.1's ``Linear`` / ``Conv2d`` forward and
  *inside* diffusers' ``DDIMScheduler.add_noise`` / ``get_velocity`` (and the ``scaled_linear`` beta schedule).  Those
  packages are absent from ``/root/reference`` and from this image (no network), and the reference ships no test or golden
  vector for them.  The stand-ins are the restatements in ``lora_ref.py`` / ``diffusion_ref.py`` of the published
  algorithms of the pinned versions (``requirements.txt:13,16``), cross-checked by closed-form identities
  (merged-weight equivalence, fp64 autograd vs closed forms, ``alpha_bar`` analytic properties) in ``tests/test_oracle.py``.
"""
