"""CPU oracle for the sampling hot path of tqch/v-diffusion-torch.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
directory; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the reported CPU baseline, never as the thing shipped.

The oracle is a from-scratch restatement (plain PyTorch fp32 on CPU for the
UNet, numpy/python fp64 for the schedule and the posterior coefficients) of

* ``v_diffusion/models/unet.py``  UNet.forward           (unet.py:286-322)
* ``v_diffusion/functions.py``    get_timestep_embedding (functions.py:11-29)
* ``v_diffusion/diffusion.py``    get_logsnr_schedule, logsnr_to_posterior[_ddim],
  p_mean_var, p_sample_step, p_sample[_progressive]      (diffusion.py:42-441)
  q_sample, from_model_out_to_pred, train_loss (mse)     (diffusion.py:242-245, 466-545)

Parity pin: ``tests/golden/make_golden.py`` imports the *unmodified reference*
from ``/root/reference`` in the build container, runs it on seeded inputs and
commits the outputs as fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those fixtures, so
the oracle is pinned to the reference itself (not "parity unpinned").
"""
from .unet_ref import unet_forward, make_state_dict, unet_config_from_json, dezero_  # noqa: F401
from .diffusion_ref import (  # noqa: F401
    logsnr_schedule, step_coefficients, p_sample, timestep_embedding, train_loss,
)
