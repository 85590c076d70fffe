"""CPU oracle for the active-learning helpers.  TEST INFRASTRUCTURE ONLY.

numpy restatement of the loss formulas in ``profit/al/aquisition_functions.py`` as pure functions of the
predictive mean / variance, and of the pick ``np.argmax(loss * mask)``.  Pinned against the unmodified reference
classes by tests/golden/make_golden_acq.py -> tests/golden/acquisition_kats.json.
"""
import numpy as np
from scipy.stats import norm

EPSILON = 1e-12          # AcquisitionFunction.EPSILON (aquisition_functions.py:38)
SIGMA_EPSILON = 1e-10    # ExpectedImprovement.SIGMA_EPSILON (:240)

norm_cdf, norm_pdf = norm.cdf, norm.pdf     # the reference's own (scipy.stats.norm, :254 / :304)


def normalize(value, minimum=None):
    """AcquisitionFunction.normalize (:76-84): column-wise min-max scaling, optional lower clamp."""
    lo, hi = value.min(axis=0), value.max(axis=0)
    out = (value - lo) / np.maximum(hi - lo, EPSILON)
    return out if minimum is None else np.maximum(out, minimum)


def loss_variance(var):
    """SimpleExploration.calculate_loss (:107-119); var (M, 1)."""
    return np.sum(var, axis=1)


def loss_distance_penalty(var, Xpred, last_point, weight):
    """ExplorationWithDistancePenalty.calculate_loss (:146-158)."""
    loss = np.sum(var, axis=1)
    scale = loss.max(axis=0)
    loss = loss + scale * (1.0 - np.exp(-weight * np.linalg.norm(Xpred - last_point, axis=1)))
    return loss / 2


def loss_weighted(mu_normalized, var, weight):
    """WeightedExploration.calculate_loss (:183-198); mu_normalized = normalize(mu) from find_next_candidates (:200)."""
    return np.sum(weight * mu_normalized + (1 - weight) * normalize(var), axis=1)


def loss_poi(mu, var, noise, ymax):
    """ProbabilityOfImprovement.calculate_loss (:212-220) with the standard-normal draws passed in."""
    return np.sum(np.maximum(mu + np.sqrt(var) * noise - ymax, 0), axis=1)


def ei_mu_part(mu, y_observed, xi, find_min):
    """ExpectedImprovement.mu_part (:262-269)."""
    mu = -mu if find_min else mu
    improvement = np.maximum(mu - np.nanmax(y_observed, axis=0), 0)
    return normalize(improvement) - xi


def loss_ei(improvement, var):
    """ExpectedImprovement.calculate_loss + sigma_part (:253-260, :271-276)."""
    sigma = normalize(np.sqrt(var), minimum=SIGMA_EPSILON)
    z = improvement / sigma
    return np.sum(improvement * norm_cdf(z) + sigma * norm_pdf(z), axis=1)


def loss_ei2(mu, var, y_observed, xi, find_min):
    """ExpectedImprovement2.calculate_loss (:303-315)."""
    mu = -mu if find_min else mu
    sigma = np.maximum(np.sqrt(var), 1e-10)
    improvement = mu - np.nanmax(y_observed, axis=0) - xi
    z = improvement / sigma
    ei = improvement * norm_cdf(z) + sigma * norm_pdf(z)
    return np.sum(ei, axis=1)


def pick(loss, visited=()):
    """idx = np.argmax(loss * mask) of _find_next_candidates (:63-68)."""
    mask = np.ones(loss.shape[0])
    mask[list(visited)] = 0
    return int(np.argmax(loss * mask))
