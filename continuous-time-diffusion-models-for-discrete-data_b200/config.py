"""Attribute-dict config (stand-in for ml_collections.ConfigDict, which the reference uses and this image lacks).

Reads the same keys as the reference: data.S, data.shape, model.{concat_dim,Q_sigma,rate_sigma,time_exp,time_base,
rate_const,t_func}, training.max_t, sampler.{name,num_steps,min_t,eps_ratio,initial_dist,num_corrector_steps,
corrector_step_size_multiplier,corrector_entry_time,is_ordinal,condition_dim,reject_multiple_jumps},
loss.{name,eps_ratio,nll_weight,min_time,one_forward_pass,logit_type,loss_type,ce_coeff}, device.
"""


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def make_config(**sections) -> ConfigDict:
    c = ConfigDict()
    for k, v in sections.items():
        c[k] = ConfigDict(v) if isinstance(v, dict) else v
    return c
