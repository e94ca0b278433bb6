"""Optimiser loops around the drop-in circuits (the callers of the hot path).

Mirrors the class surface of the reference's `optimization.py` (`McCleanOpt`, `QaoaOpt` at
optimization.py:41-129 and the update rules `Adam`, `GradientDescent`, `RateDecayOnPlateau` at
:131-194): same constructor arguments, attributes (`cost_history`, `param_history`, `iter`,
`optimizer`, `circuit`) and step semantics, so scripts written against the reference run unchanged.
Every `step()` is one `grad_run` (or sampled-gradient) call on the GPU circuit plus an O(L*n)
host update; parity with the reference's loops is pinned by tests/golden/gv13_optimizers.npz.
"""
import numpy as np


# ---------------------------------------------------------------------------------------------
# update rules (optimization.py:131-194)
# ---------------------------------------------------------------------------------------------
class _Rule:
    """Holds the parameter array (updated in place, like the reference) and the step counter."""

    def __init__(self, parameters, hyper_parameters):
        self.parameters = parameters
        self.step_size = hyper_parameters.get('step_size', 1e-3)
        self.iter = 0


GradientDescentOptimizer = _Rule      # the reference's name of the base class (optimization.py:131)


class Adam(_Rule):
    def __init__(self, parameters, hyper_parameters):
        super().__init__(parameters, hyper_parameters)
        self.beta1 = hyper_parameters.get('beta1', 0.9)
        self.beta2 = hyper_parameters.get('beta2', 0.999)
        self.eps = hyper_parameters.get('eps', 1e-8)
        shape = np.shape(parameters)
        self.m, self.v = np.zeros(shape), np.zeros(shape)
        self.m_hat, self.v_hat = np.zeros(shape), np.zeros(shape)

    def step(self, gradient, *_):
        self.iter += 1
        b1, b2 = self.beta1, self.beta2
        self.m = b1 * self.m + (1 - b1) * gradient
        self.v = b2 * self.v + (1 - b2) * gradient ** 2
        self.m_hat = self.m / (1 - b1 ** self.iter)       # bias-corrected moments
        self.v_hat = self.v / (1 - b2 ** self.iter)
        self.parameters -= self.step_size * self.m_hat / (np.sqrt(self.v_hat) + self.eps)


class GradientDescent(_Rule):
    def __init__(self, parameters, hyper_parameters):
        super().__init__(parameters, hyper_parameters)
        self.decay_function = hyper_parameters.get('decay_function', lambda step_size, it: step_size)
        self.constant_step = 'decay_function' not in hyper_parameters

    def step(self, gradient, *_):
        self.iter += 1
        self.parameters -= self.decay_function(self.step_size, self.iter) * gradient


class RateDecayOnPlateau(_Rule):
    def __init__(self, parameters, hyper_parameters):
        super().__init__(parameters, hyper_parameters)
        self.plateau_length = hyper_parameters.get('plateau_length', 10)
        self.decay_rate = hyper_parameters.get('decay_rate', 0.5)
        self.plateau_counter = 0
        self.cost = 1e10

    def step(self, gradient, new_cost):
        self.iter += 1
        if new_cost > self.cost:             # no improvement: count towards a plateau, shrink the rate when it is reached
            self.plateau_counter += 1
            if self.plateau_counter >= self.plateau_length:
                self.step_size *= self.decay_rate
                self.plateau_counter = 0
        else:
            self.cost = new_cost
            self.plateau_counter = 0
        self.parameters -= self.step_size * gradient


_RULES = {'Adam': Adam, 'GradientDescent': GradientDescent, 'RateDecayOnPlateau': RateDecayOnPlateau}


# ---------------------------------------------------------------------------------------------
# circuit-specific loops (optimization.py:3-129)
# ---------------------------------------------------------------------------------------------
class ParametrizedCircuitOptimizer:
    """Base class: tracks the circuit, the iteration counter and the chosen update rule."""

    def init(self, circuit, max_iter):
        self.circuit = circuit
        self.max_iter = max_iter
        self.iter = 0

    def step(self):
        pass

    def reset(self):
        pass

    def pick(self, optimizer, ini_parameters):
        try:
            rule = _RULES[optimizer['name']]
        except KeyError:
            raise ValueError('No optimizer {} known.'.format(optimizer))     # optimization.py:39
        self.optimizer = rule(ini_parameters, optimizer)

    def __str__(self):
        return str(self.optimizer_info)

    def _check_budget(self):
        if self.iter >= self.max_iter:
            print('Maximum amount of iterations reached: {}.'.format(self.max_iter))


class McCleanOpt(ParametrizedCircuitOptimizer):
    """Optimiser loop for `McClean` circuits: `step()` = one gradient evaluation + one update of `circuit.angles`."""

    def __init__(self, circuit, optimizer, max_iter=1000, **kwargs):
        self.init(circuit, max_iter)
        self.param_history = np.zeros([max_iter, circuit.lnum, circuit.qnum], dtype='double')
        self.cost_history = np.zeros(max_iter, dtype='double')
        ini_parameters = kwargs.get('ini_parameters', circuit.angles)
        self.param_history[0] = ini_parameters
        self.circuit.angles = ini_parameters          # shared with the update rule, which edits it in place
        self.optimizer_info = optimizer
        self.pick(optimizer, ini_parameters)

    def step(self, shot_num=0, dense_mode=True, component_sampling=False):
        self._check_budget()
        c = self.circuit
        if shot_num == 0:
            e, g = c.grad_run_with_component_sampling() if component_sampling else c.grad_run()
        elif dense_mode:
            fn = c.sample_grad_dense_with_component_sampling if component_sampling else c.sample_grad_dense
            e, g = fn(shot_num=shot_num)
        else:
            fn = c.sample_grad_with_component_sampling if component_sampling else c.sample_grad
            e, g = fn(shot_num=shot_num)
        self.cost_history[self.iter] = e
        self.optimizer.step(g, e)
        self.iter += 1
        self.param_history[self.iter] = self.optimizer.parameters
        self.circuit.angles = self.optimizer.parameters

    def run(self, steps):
        """`steps` exact-gradient steps (the same sequence as `steps` calls of `step()`), executed as ONE device-side
        loop when the circuit and the update rule support it (`McClean.optimize_on_device`: Adam, GradientDescent with a
        constant step, RateDecayOnPlateau): gradient, parameter update and the next gate tables stay on the GPU, the
        histories come back once at the end.  Falls back to `step()` otherwise."""
        opt = self.optimizer
        steps = int(min(steps, self.max_iter - 1 - self.iter))
        rule = {Adam: 0, GradientDescent: 1, RateDecayOnPlateau: 2}.get(type(opt))
        device_ok = (steps > 0 and rule is not None and hasattr(self.circuit, 'optimize_on_device')
                     and self.circuit.qnum >= 4 and (rule != 1 or opt.constant_step))
        if not device_ok:
            for _ in range(max(steps, 0)):
                self.step()
            return
        hyper = np.array([opt.step_size, getattr(opt, 'beta1', 0.), getattr(opt, 'beta2', 0.), getattr(opt, 'eps', 0.),
                          getattr(opt, 'plateau_length', 0), getattr(opt, 'decay_rate', 0.), getattr(opt, 'cost', 0.),
                          getattr(opt, 'plateau_counter', 0)], dtype=np.float64)
        self.circuit.angles = np.ascontiguousarray(opt.parameters, dtype=np.float64)
        m = opt.m if rule == 0 else None
        v = opt.v if rule == 0 else None
        cost, hist, it, hyper = self.circuit.optimize_on_device(rule, hyper, opt.iter, steps, m=m, v=v)
        self.cost_history[self.iter:self.iter + steps] = cost
        self.param_history[self.iter + 1:self.iter + 1 + steps] = hist
        self.iter += steps
        opt.iter = it
        opt.step_size = hyper[0]
        if rule == 0:
            opt.m_hat = opt.m / (1 - opt.beta1 ** opt.iter)
            opt.v_hat = opt.v / (1 - opt.beta2 ** opt.iter)
        if rule == 2:
            opt.cost, opt.plateau_counter = hyper[6], int(hyper[7])
        opt.parameters[...] = self.circuit.angles
        self.circuit.angles = opt.parameters

    def reset(self, **kwargs):
        self.__init__(self.circuit, kwargs.get('optimizer', self.optimizer_info), kwargs.get('max_iter', self.max_iter),
                      ini_parameters=kwargs.get('ini_parameters', self.param_history[0]))


class QaoaOpt(ParametrizedCircuitOptimizer):
    """Optimiser loop for `Qaoa` circuits; parameters are rows (beta_i, gamma_i)."""

    def __init__(self, circuit, optimizer, betas, gammas, max_iter=1000):
        self.init(circuit, max_iter)
        self.param_history = np.zeros([max_iter, circuit.lnum, 2], dtype='double')
        self.cost_history = np.zeros(max_iter, dtype='double')
        start = np.array([betas, gammas]).transpose()
        self.param_history[0] = start
        self.optimizer_info = optimizer
        self.pick(optimizer, start)

    def step(self, shot_num=0, dense_mode=True):
        self._check_budget()
        betas, gammas = self.param_history[self.iter, :, 0], self.param_history[self.iter, :, 1]
        if shot_num == 0:
            e, g = self.circuit.grad_run(betas, gammas)
        elif dense_mode:
            e, g = self.circuit.sample_grad_dense(betas, gammas, shot_num=shot_num)
        else:
            raise ValueError('dense_mode must be True, sparse sampling method is not implmented yet.')   # optimization.py:121
        self.cost_history[self.iter] = e
        self.optimizer.step(g, e)
        self.iter += 1
        self.param_history[self.iter] = self.optimizer.parameters

    def run(self, steps):
        """`steps` exact-gradient steps as ONE device-side loop (`Qaoa.optimize_on_device`, qr_qaoa_optimize) where the
        update rule and the Hamiltonian allow it (Adam, constant-step GradientDescent, RateDecayOnPlateau; integer-valued
        H such as MaxCut); the same sequence as `steps` calls of `step()`, which is also the fallback."""
        opt = self.optimizer
        steps = int(min(steps, self.max_iter - 1 - self.iter))
        rule = {Adam: 0, GradientDescent: 1, RateDecayOnPlateau: 2}.get(type(opt))
        device_ok = (steps > 0 and rule is not None and hasattr(self.circuit, 'optimize_on_device')
                     and self.circuit.qnum >= 4 and (rule != 1 or opt.constant_step))
        if device_ok:
            hyper = np.array([opt.step_size, getattr(opt, 'beta1', 0.), getattr(opt, 'beta2', 0.), getattr(opt, 'eps', 0.),
                              getattr(opt, 'plateau_length', 0), getattr(opt, 'decay_rate', 0.), getattr(opt, 'cost', 0.),
                              getattr(opt, 'plateau_counter', 0)], dtype=np.float64)
            params = np.ascontiguousarray(opt.parameters, dtype=np.float64)
            m = opt.m if rule == 0 else None
            v = opt.v if rule == 0 else None
            try:
                cost, hist, it, hyper = self.circuit.optimize_on_device(rule, hyper, opt.iter, steps, params, m=m, v=v)
            except ValueError as exc:
                if 'integer-valued' not in str(exc):
                    raise
                device_ok = False          # a Hamiltonian with non-integer values: host loop below
        if not device_ok:
            for _ in range(max(steps, 0)):
                self.step()
            return
        self.cost_history[self.iter:self.iter + steps] = cost
        self.param_history[self.iter + 1:self.iter + 1 + steps] = hist
        self.iter += steps
        opt.iter = it
        opt.step_size = hyper[0]
        if rule == 0:
            opt.m_hat = opt.m / (1 - opt.beta1 ** opt.iter)
            opt.v_hat = opt.v / (1 - opt.beta2 ** opt.iter)
        if rule == 2:
            opt.cost, opt.plateau_counter = hyper[6], int(hyper[7])
        opt.parameters[...] = params

    def reset(self, **kwargs):
        if 'betas' in kwargs and 'gammas' in kwargs:
            self.param_history[0] = np.array([kwargs['betas'], kwargs['gammas']]).transpose()
        self.__init__(self.circuit, kwargs.get('optimizer', self.optimizer_info), self.param_history[0, :, 0],
                      self.param_history[0, :, 1], kwargs.get('max_iter', self.max_iter))
