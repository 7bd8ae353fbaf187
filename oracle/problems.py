"""Oracle-side builder for the synthetic problem dicts of workloads.make_problem (test infrastructure: used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only)."""


def build_oracle(prob, faithful=False):
    """Instantiate oracle/reference_dgp.py for a problem dict."""
    import torch
    from oracle import reference_dgp as R
    R.settings.jitter = prob['jitter']
    R.SVGP_Layer.faithful = faithful
    kcls = R.RBF if prob['kern'] == 'rbf' else R.Matern52
    mfs = {'zero': lambda l: R.Zero(), 'identity': lambda l: R.Identity(), 'linear': lambda l: R.Linear(l['W'])}
    layers = []
    for lay in prob['layers']:
        kern = kcls(lay['din'], variance=lay['var'], lengthscales=lay['ls'])
        if lay.get('wvar') is not None:
            kern = R.Sum([kern, R.White(lay['din'], variance=lay['wvar'])])
        layer = R.SVGP_Layer(kern, lay['Z'], lay['dout'], mfs[lay['mean']](lay), white=lay['white'],
                             input_prop_dim=lay.get('ipd'))
        layer.q_mu = torch.as_tensor(lay['q_mu']).clone()
        layer.q_sqrt = torch.as_tensor(lay['q_sqrt']).clone()
        layers.append(layer)
    if prob.get('lik') == 'bernoulli':
        lik = R.Bernoulli()
    else:
        lik = R.MultiClass(prob['n_classes']) if prob['n_classes'] else R.Gaussian(prob['lik_var'])
    return R.DGP_Base(prob['X'], prob['Y'], lik, layers, num_samples=prob['S'], num_data=prob['num_data'])
