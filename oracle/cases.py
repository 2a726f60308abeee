"""TEST INFRASTRUCTURE ONLY — the named parity cases shared by oracle/make_golden.py and tests/.

Each case fixes a forward process, a stub score network (StubNet: bitwise identical on CPU and CUDA), the sampler
config and a Philox seed, at sizes the CPU oracle finishes in seconds.
"""
from __future__ import annotations

FORWARD = {
    "gauss256": dict(mixin="GaussianTargetRate", kind="gaussian", S=256,
                     model=dict(rate_sigma=6.0, Q_sigma=512.0, time_exp=100.0, time_base=3.0)),
    "gauss32": dict(mixin="GaussianTargetRate", kind="gaussian", S=32,
                    model=dict(rate_sigma=3.0, Q_sigma=20.0, time_exp=50.0, time_base=1.0)),
    "uni2": dict(mixin="UniformRate", kind="uniform", S=2, model=dict(rate_const=1.0)),
    "univar3_logsqr": dict(mixin="UniformVariantRate", kind="uniform_variant", S=3, model=dict(rate_const=2.0, t_func="log_sqr")),
    "univar2_sqrtcos": dict(mixin="UniformVariantRate", kind="uniform_variant", S=2, model=dict(rate_const=2.0, t_func="sqrt_cos")),
    "univar5_log": dict(mixin="UniformVariantRate", kind="uniform_variant", S=5,
                        model=dict(rate_const=0.5, t_func="log", time_base=1.0, time_exp=10.0)),
    "bd16": dict(mixin="BirthDeathForwardBase", kind="birth_death", S=16, model=dict(sigma_min=1.0, sigma_max=10.0)),
}
FORWARD_TIMES = {"gauss256": [0.05, 1.0], "gauss32": [0.01, 0.2, 0.6, 1.0], "uni2": [0.01, 0.5, 1.0],
                 "univar3_logsqr": [0.001, 0.3, 1.0], "univar2_sqrtcos": [0.007, 0.5, 0.99999],
                 "univar5_log": [0.01, 0.5, 1.0], "bd16": [0.01, 0.5, 1.0]}

# (name, forward, N, D, loss.name, logit_type, stub(scale, width), t)
RATES = [
    ("rr_gauss256_tauldr", "gauss256", 2, 5, "CTElbo", None, (0.5, 10.0), 0.3),
    ("rr_gauss256_revprob", "gauss256", 2, 5, "CatRM", "reverse_prob", (0.5, 10.0), 0.3),
    ("rr_gauss32_tauldr", "gauss32", 3, 7, "NLL", None, (1.0, 3.0), 0.05),
    ("rr_gauss32_direct", "gauss32", 3, 7, "CatRM", "direct", (1.0, 3.0), 0.5),
    ("rr_gauss32_logscale", "gauss32", 3, 7, "ScoreElbo", "reverse_logscale", (1.0, 3.0), 0.5),
    ("rr_univar3_tauldr", "univar3_logsqr", 4, 9, "CTElboLambda", None, (1.0, None), 0.4),
    ("rr_univar3_revprob", "univar3_logsqr", 4, 9, "CatRMNLL", "reverse_prob", (1.0, None), 0.4),
    ("rr_uni2_direct", "uni2", 4, 8, "SDDMElbo", "direct", (1.0, None), 0.7),
]

_S = dict(eps_ratio=1e-9, corrector_step_size_multiplier=1.5, corrector_entry_time=0.0, num_corrector_steps=0,
          is_ordinal=True, initial_dist="gaussian")

# (name, sampler class, forward, N, D, loss.name, logit_type, stub(scale,width), sampler overrides, max_t, seed)
SAMPLERS = [
    ("taul_gauss256_ord", "TauL", "gauss256", 8, 12, "CTElbo", None, (0.3, 12.0),
     dict(num_steps=8, min_t=0.01), 1.0, 101),
    ("taul_gauss256_corr_crm", "TauL", "gauss256", 8, 12, "CatRM", "reverse_prob", (0.3, 12.0),
     dict(num_steps=6, min_t=0.01, corrector_entry_time=0.3, num_corrector_steps=2), 1.0, 102),
    ("taul_uni2_nonord", "TauL", "univar2_sqrtcos", 64, 32, "CTElboLambda", None, (1.0, None),
     dict(num_steps=20, min_t=0.007, is_ordinal=False, initial_dist="uniform"), 0.99999, 103),
    ("taul_gauss32_direct", "TauL", "gauss32", 16, 10, "CatRM", "direct", (1.0, 3.0),
     dict(num_steps=10, min_t=0.01), 1.0, 104),
    ("lbjf_univar3_pc", "LBJF", "univar3_logsqr", 32, 225, "NLL", None, (1.0, None),
     dict(num_steps=12, min_t=0.001, corrector_entry_time=0.1, num_corrector_steps=2, initial_dist="uniform"), 1.0, 105),
    ("lbjf_gauss32_crm", "LBJF", "gauss32", 16, 10, "CatRMNLL", "reverse_prob", (1.0, 3.0),
     dict(num_steps=10, min_t=0.01), 1.0, 106),
    ("midpoint_univar3", "MidPointTauL", "univar3_logsqr", 32, 40, "CTElbo", None, (0.3, 0.7),
     dict(num_steps=10, min_t=0.001, is_ordinal=False, initial_dist="uniform"), 1.0, 107),
    ("midpoint_uni2_ord", "MidPointTauL", "univar2_sqrtcos", 32, 32, "ScoreElbo", "reverse_prob", (1.0, None),
     dict(num_steps=10, min_t=0.007, is_ordinal=True, initial_dist="uniform"), 0.99999, 108),
    ("pctaul_univar3", "PCTauL", "univar3_logsqr", 16, 30, "CatRM", "direct", (0.3, 0.7),
     dict(num_steps=10, min_t=0.001, corrector_entry_time=0.3, num_corrector_steps=1, initial_dist="uniform"), 1.0, 109),
    ("pctaul_gauss32", "PCTauL", "gauss32", 8, 10, "CTElbo", None, (1.0, 3.0),
     dict(num_steps=8, min_t=0.01, corrector_entry_time=0.2, num_corrector_steps=1), 1.0, 110),
    ("condtaul_gauss32", "ConditionalTauLeaping", "gauss32", 8, 10, "CTElbo", None, (1.0, 3.0),
     dict(num_steps=8, min_t=0.01, condition_dim=4, reject_multiple_jumps=True), 1.0, 111),
    ("condpctaul_gauss32", "ConditionalPCTauLeaping", "gauss32", 8, 10, "CTElbo", None, (1.0, 3.0),
     dict(num_steps=8, min_t=0.01, condition_dim=4, reject_multiple_jumps=True, corrector_entry_time=0.3,
          num_corrector_steps=1), 1.0, 112),
    ("exact_univar3", "ExactSampling", "univar3_logsqr", 32, 40, "CTElbo", None, (0.3, 0.7),
     dict(num_steps=10, min_t=0.001, initial_dist="uniform"), 1.0, 113),
    ("exact_uni2", "ExactSampling", "uni2", 32, 32, "CatRM", "direct", (1.0, None),
     dict(num_steps=12, min_t=0.01, initial_dist="uniform"), 1.0, 114),
    ("exact_gauss32", "ExactSampling", "gauss32", 8, 10, "CTElbo", None, (1.0, 3.0),
     dict(num_steps=8, min_t=0.01), 1.0, 115),
]


# Whole-sampler cases at S = 256 for the classes whose reference implementation cannot run there (MidPointTauL has no
# state_change table for DiscreteCIFAR10) or was fixture-tested at small S only (LBJF): checked against
# the oracle samplers, which are pinned to the reference at S = 2 / 3 / 32 by the fixtures above.  BASELINE config C5.
SAMPLERS_S256 = [
    ("midpoint_gauss256_sddm", "MidPointTauL", "gauss256", 16, 24, "SDDMElbo", "reverse_prob", (0.3, 12.0),
     dict(num_steps=6, min_t=0.01), 1.0, 131),
    ("midpoint_gauss256_tauldr", "MidPointTauL", "gauss256", 16, 24, "CTElbo", None, (0.3, 12.0),
     dict(num_steps=6, min_t=0.01), 1.0, 132),
    ("lbjf_gauss256", "LBJF", "gauss256", 16, 24, "CTElbo", None, (0.3, 12.0),
     dict(num_steps=6, min_t=0.01, corrector_entry_time=0.3, num_corrector_steps=1), 1.0, 133),
]


def sampler_cfg(make_cfg, case):
    """Build the config object (same keys as the reference's ml_collections configs) for a SAMPLERS case."""
    name, cls, fwd, N, D, loss_name, logit_type, stub, over, max_t, seed = case
    f = FORWARD[fwd]
    S = f["S"]
    s = dict(_S)
    s.update(over)
    s["name"] = cls
    data_name = {2: "SyntheticData", 3: "Maze3S"}.get(S, "DiscreteCIFAR10")
    loss = dict(name=loss_name, eps_ratio=1e-9, nll_weight=0.001, min_time=0.01, one_forward_pass=True,
                logit_type=logit_type if logit_type else "reverse_prob", loss_type="rm", ce_coeff=0.0)
    model = dict(f["model"])
    model["concat_dim"] = D
    model.setdefault("Q_sigma", 20.0)
    model["log_prob"] = "cat"
    return make_cfg(data=dict(S=S, shape=[D], name=data_name), model=model, training=dict(max_t=max_t, n_iters=1000),
                    sampler=s, loss=loss, device="cpu")


# (name, loss class, forward, B, D, loss overrides, t_hi, seed, n_iter)
# t_hi is the upper end of the reference's time draw for that class (losses.py: max_t / 1.0).
LOSSES = [
    ("ctelbo_gauss32", "CTElbo", "gauss32", 6, 10, dict(), 1.0, 201, 0),
    ("ctelbo_gauss32_2pass", "CTElbo", "gauss32", 6, 10, dict(one_forward_pass=False), 1.0, 202, 0),
    ("ctelbo_gauss256", "CTElbo", "gauss256", 3, 6, dict(nll_weight=0.01), 1.0, 203, 0),
    ("nll_univar3", "NLL", "univar3_logsqr", 8, 9, dict(), 1.0, 204, 0),
    ("ctelbolambda_uni2", "CTElboLambda", "univar2_sqrtcos", 8, 16, dict(), 0.99999, 205, 300),
    ("condctelbo_gauss32", "CondCTElbo", "gauss32", 6, 10, dict(condition_dim=4), 1.0, 206, 0),
    ("condctelbo_gauss32_2pass", "CondCTElbo", "gauss32", 6, 10, dict(condition_dim=4, one_forward_pass=False), 1.0, 217, 0),
    ("catrm_gauss32_rm", "CatRM", "gauss32", 6, 10, dict(logit_type="reverse_prob", loss_type="rm"), 1.0, 207, 0),
    ("catrm_gauss32_mle_direct", "CatRM", "gauss32", 6, 10, dict(logit_type="direct", loss_type="mle"), 1.0, 208, 0),
    ("catrm_univar3_elbo", "CatRM", "univar3_logsqr", 8, 9, dict(logit_type="reverse_prob", loss_type="elbo", ce_coeff=0.25), 1.0, 209, 0),
    ("catrm_gauss32_logscale", "CatRM", "gauss32", 4, 6, dict(logit_type="reverse_logscale", loss_type="rm"), 1.0, 210, 0),
    ("catrmnll_gauss256", "CatRMNLL", "gauss256", 3, 6, dict(logit_type="reverse_prob", loss_type="rm", nll_weight=0.01), 1.0, 211, 0),
    ("scoreelbo_gauss32", "ScoreElbo", "gauss32", 6, 10, dict(logit_type="reverse_prob"), 1.0, 212, 0),
    ("sddmelbo_gauss256", "SDDMElbo", "gauss256", 3, 6, dict(logit_type="reverse_prob", nll_weight=0.01), 1.0, 213, 0),
    ("sddmelbo_uni2_direct", "SDDMElbo", "uni2", 8, 16, dict(logit_type="direct"), 1.0, 214, 0),
    ("scoreelbo_gauss32_logscale", "ScoreElbo", "gauss32", 4, 6, dict(logit_type="reverse_logscale"), 1.0, 215, 0),
    ("nlloriginal_gauss32", "NLLOriginal", "gauss32", 6, 10, dict(), 1.0, 216, 0),
]

# losses whose calc_loss takes (minibatch, state) rather than (state, minibatch)  
LOSS_MINIBATCH_FIRST = ("SDDMElbo", "CondCTElbo", "CatRMNLL", "ScoreElbo")


def loss_cfg(make_cfg, case, device="cpu"):
    name, cls, fwd, B, D, over, t_hi, seed, n_iter = case
    f = FORWARD[fwd]
    S = f["S"]
    loss = dict(name=cls, eps_ratio=1e-9, nll_weight=0.001, min_time=0.01, one_forward_pass=True,
                logit_type="reverse_prob", loss_type="rm", ce_coeff=0.0, condition_dim=0)
    loss.update(over)
    model = dict(f["model"])
    model["concat_dim"] = D
    model.setdefault("Q_sigma", 20.0)
    return make_cfg(data=dict(S=S, shape=[D], name="DiscreteCIFAR10"), model=model,
                    training=dict(max_t=t_hi if cls in ("CTElbo", "NLL", "CTElboLambda", "CatRMNLL") else 1.0, n_iters=1000),
                    loss=loss, device=device)
