"""B200-native drop-in for the factor-update path of nn-fac (module paths mirror the reference).

    import nn_fac.nmf as nmf                       # nmf.nmf / compute_nmf / one_nmf_step
    import nn_fac.ntf as ntf                       # ntf.ntf / compute_ntf / one_ntf_step
    import nn_fac.ntd as ntd                       # ntd.ntd / compute_ntd / one_ntd_step_mu
    import nn_fac.update_rules.nnls as nnls        # hals_nnls_acc
    import nn_fac.update_rules.mu as mu            # switch_alternate_mu / mu_betadivmin / mu_tensorial
    import nn_fac.utils.beta_divergence as beta_div

Every operator runs in libnnfac_b200.so (hand-written sm_100a CUDA); there is no CPU fallback.
Like the reference's own nn_fac/__init__.py, nothing is re-exported here.
"""
