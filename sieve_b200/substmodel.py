"""Host-side model pieces that stay outside the device path: SIEVE's fixed 4-state genotype rate matrix,
the transition-probability matrices handed to the core, and the gamma site-rate categories.

The reference keeps these on the host as well (SURVEY.md section 8 rows a7 / boundary inputs): the 16-double P matrices
are produced by BEAST's GeneralSubstitutionModel eigen code (external) from the Q below and given to
`setNodeMatrix` (likelihood/ScsTreeLikelihood.java:567-578).
"""
import numpy as np

N_STATES = 4          # 0/0, 0/1, 1/1, 1/1'   (substitutionmodel/ScsFiniteMuExtendedModel.java:31-34)
ROOT_GENOTYPE = 0     # :37
CONST_GENOTYPE = 0    # :40


def rate_matrix():
    """substitutionmodel/ScsFiniteMuExtendedModel.java:62-86 (relative rates) normalised as
    substitutionmodel/ScsSubstitutionModelBase.java:200-231 (sum of -diag == 1)."""
    rel = np.array([
        [0.0, 1.0, 0.0, 0.0],
        [1.0 / 6, 0.0, 1.0 / 6, 1.0 / 3],
        [0.0, 1.0 / 3, 0.0, 2.0 / 3],
        [0.0, 1.0 / 3, 1.0 / 3, 0.0],
    ])
    Q = rel.copy()
    for i in range(4):
        Q[i, i] = -rel[i].sum()
    subst = -np.trace(Q)
    return Q / subst


# Spectral form of exp(Q d): Q has eigenvalues 0, -0.2, -0.4 (twice) and is diagonalisable, so
#   P(d) = I + expm1(-0.2 d) * PI1 + expm1(-0.4 d) * PI2
# with the two projectors below (exact rationals).  expm1 keeps the O(d) entries accurate for short branches.
PI1 = np.array([[3, 6, -3, -6], [1, 2, -1, -2], [-1, -2, 1, 2], [-1, -2, 1, 2]], dtype=np.float64) / 8.0
PI2 = np.array([[9, -18, 3, 6], [-3, 6, -1, -2], [1, -2, 11, -10], [1, -2, -5, 6]], dtype=np.float64) / 16.0


def transition_probabilities(distance):
    """P = exp(Q * distance), row = parent state, column = child state (the layout `setNodeMatrix` takes,
    likelihood/ScsBeerLikelihoodCore.java:107-108). `distance` = branch length * branch rate * site rate."""
    if distance == 0.0:
        return np.eye(4)
    P = np.eye(4) + np.expm1(-0.2 * distance) * PI1 + np.expm1(-0.4 * distance) * PI2
    return np.clip(P, 0.0, None)


def gamma_category_rates(shape, k=4, mu=1.0):
    """BEAST1-style discrete gamma as ScsSiteModel forces it (sitemodel/ScsSiteModel.java:16): medians of k
    equiprobable bins, normalised to mean 1, times the mutation rate. Proportions are all 1/k."""
    from scipy.stats import gamma
    q = (2.0 * np.arange(k) + 1.0) / (2.0 * k)
    r = gamma.ppf(q, a=shape, scale=1.0 / shape)
    r = r / r.mean()
    return r * mu, np.full(k, 1.0 / k)


def node_distances(tree, cat_rates, branch_rates=None, clock_rate=1.0):
    """[2n][K] distances = length * branchRate * siteRate_k (the arguments of getTransitionProbabilities at
    likelihood/ScsTreeLikelihood.java:567-571); the root's row is unused."""
    bl = tree.branch_lengths()
    br = clock_rate * (np.ones(tree.node_count) if branch_rates is None else np.asarray(branch_rates))
    return (bl * br)[:, None] * np.asarray(cat_rates)[None, :]


def node_matrices(tree, cat_rates, branch_rates=None, clock_rate=1.0):
    """[2n][K][16] transition matrices for every node (row of the root unused), as the traverse would set them
    (likelihood/ScsTreeLikelihood.java:557-578): distance = length * branchRate * siteRate_k."""
    n_nodes = tree.node_count
    K = len(cat_rates)
    out = np.zeros((n_nodes, K, 16))
    bl = tree.branch_lengths()
    for i in range(n_nodes):
        if i == tree.root:
            out[i] = np.eye(4).reshape(16)
            continue
        br = clock_rate * (1.0 if branch_rates is None else branch_rates[i])
        for k in range(K):
            out[i, k] = transition_probabilities(bl[i] * br * cat_rates[k]).reshape(16)
    return out
