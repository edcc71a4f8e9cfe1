"""CPU/PyTorch-fp32 ORACLE of the MakeupDiffuse denoising hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain restatement of the algorithm the reference executes on the path
``MKDDIMSampler.{sample,reconstruct,denoising_step}`` -> ``apply_model`` -> ControlNet + ControlledUnetModel.
It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package ``makeupdiffuse_b200``
never imports it and has no CPU fallback.

PARITY PINNED FOR THE SAMPLER, UNPINNED FOR THE NETWORKS.  The reference ships no tests, fixtures or golden
vectors for this path, and the network arithmetic lives in the un-vendored, un-pinned
third-party lllyasviel/ControlNet packages ``ldm`` / ``cldm`` which are not importable here.  What CAN run here is
the reference's own ``diffmk/cddim.py``: ``tests/golden/make_golden_ref_sampler.py`` executes that file unmodified
(with a stand-in for the one upstream module it star-imports) over a closed-form toy denoiser, and
``tests/test_ref_sampler_golden.py`` holds ``oracle.MKDDIMSampler.{denoising_step, reconstruct}`` to those outputs:
bit-identical on all 11 cases (CFG on dict / list / tensor conditioning, eta > 0 with its RNG consumption,
repeat_noise, truncated loops, use_original_steps).  For everything else the oracle follows

  * in-repo, cite-able code: ``diffmk/cddim.py:9-100`` (per-step DDIM math, CFG batching order, loop/index
    convention), ``diffmk/makeup_diffuse.py:152-170`` (ControlNet -> x control_scales -> UNet dataflow),
    ``diffmodels/base_diffusion_makeup.yaml:4-8,41-50,52-84`` (every hyper-parameter),
    ``runs/train.py:60-62`` (hint conv weight shape / key name);
  * the published algorithm of upstream ControlNet (module trees, op order, schedule formulas),
    restated from its public description and pinned by the structural / known-answer tests K1-K7 of
    SURVEY.md §8(c) (parameter counts 859 520 964 / 361 279 552, state-dict key names, schedule values,
    the eps==0 closed form, CFG identities, ...), which live in ``tests/test_oracle_kat.py``.
"""
from .init import hash_uniform, seeded_state_dict  # noqa: F401
from .nets import ControlNet, ControlledUnetModel, timestep_embedding  # noqa: F401
from .ldm import OracleControlLDM, linear_beta_alphas_cumprod  # noqa: F401
from .ddim import DDIMSampler, MKDDIMSampler  # noqa: F401
