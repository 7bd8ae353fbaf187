"""gpflow.likelihoods stand-ins (reference call sites utils.py:88-121; Gaussian, MultiClass, Bernoulli): descriptors only.  variational_expectations
(the training path) and the prediction epilogues predict_mean_and_var / predict_density all run on the device
(csrc/lik_adam.cu: k_lik_*, k_predict_y_*, k_density_*) and are reached through the model: compute_log_likelihood,
predict_y, predict_density (dgp.py:92-126)."""
from .params import Parameter, Parameterized


class Likelihood(Parameterized):
    code = -1


class Gaussian(Likelihood):
    code = 0

    def __init__(self, variance=1.0):
        self.variance = Parameter(variance)


class MultiClass(Likelihood):
    """MultiClass(K) with the default RobustMax(epsilon=1e-3) link (demos/demo_mnist.ipynb:102)."""
    code = 1

    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        if epsilon != 1e-3:
            raise NotImplementedError("RobustMax epsilon is fixed at 1e-3 in the device kernel")
        self.epsilon = float(epsilon)


class Bernoulli(Likelihood):
    """gpflow.likelihoods.Bernoulli() with its default probit link (tests/test_dgp.py:48-54): 20-point Gauss-Hermite
    variational expectations, closed-form predict_mean_and_var (csrc/lik_adam.cu k_lik_bernoulli, k_predict_y_bernoulli)."""
    code = 2
