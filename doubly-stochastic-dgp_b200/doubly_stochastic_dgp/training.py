"""gpflow.training / gpflow.actions stand-ins with the call shapes the reference uses:

    AdamOptimizer(0.01).minimize(model, maxiter=iterations)                        demos/run_regression.py:83
    NatGradOptimizer(gamma=1.).minimize(m, var_list=[[q_mu, q_sqrt]], maxiter=1)   tests/test_collapsed.py:99-100
    ng_action = NatGradOptimizer(gamma).make_optimize_action(model, var_list=ng_vars)
    adam_action = AdamOptimizer(0.001).make_optimize_action(model)
    Loop([ng_action, adam_action], stop=iterations)()                              demos/using_natural_gradients.ipynb

Each action is one device step (dsdgp_train_step / dsdgp_natgrad_step): minibatch draw, ELBO forward, backward, update.
As with GPflow, every action computes its own gradient on its own minibatch."""


class AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.hyper = (float(learning_rate), float(beta1), float(beta2), float(epsilon))

    def make_optimize_action(self, model, var_list=None):
        if var_list is not None:
            raise NotImplementedError("AdamOptimizer(var_list=...): use param.set_trainable(False) on the others")
        if model._adam != self.hyper:
            model.adam_init(*self.hyper)
        return model.train_step

    def minimize(self, model, maxiter=1000, var_list=None):
        act = self.make_optimize_action(model, var_list=var_list)
        e = None
        for _ in range(int(maxiter)):
            e = act()
        return e


class NatGradOptimizer:
    def __init__(self, gamma):
        self.gamma = float(gamma)

    def make_optimize_action(self, model, var_list=None):
        model._natgrad_layers(var_list)      # validates now, like GPflow building the op
        return lambda: model.natgrad_step(var_list=var_list, gamma=self.gamma)

    def minimize(self, model, var_list=None, maxiter=1000):
        act = self.make_optimize_action(model, var_list=var_list)
        e = None
        for _ in range(int(maxiter)):
            e = act()
        return e


class Loop:
    """gpflow.actions.Loop([actions], stop=n)(): run the actions in order, n times."""
    def __init__(self, actions, stop=1):
        self.actions = list(actions) if isinstance(actions, (list, tuple)) else [actions]
        self.stop = int(stop)

    def __call__(self):
        out = None
        for _ in range(self.stop):
            for a in self.actions:
                out = a()
        return out
