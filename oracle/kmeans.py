"""Restatement of models/Kmeans_2.py (in-graph multi-try hard/soft k-means).

Test infrastructure (see oracle/__init__.py).  Reproduces the reference quirks
listed in SURVEY.md section 8(a16):
  * silent bins get distance 0 to every centroid -> label 0, and are counted in
    cluster 0's denominator with zero vectors (Kmeans_2.py:148, 158-165);
  * ``notsilent`` is tiled try-major (Kmeans_2.py:80) while X is tiled
    batch-major (:48-52): row r of the [b*tries] axis uses mask row (r mod b);
  * empty cluster -> 0/0 = NaN (callers must avoid it; tests exclude it);
  * the random initial rows come from a host callback (Kmeans_2.py:61-65); here
    the caller supplies them (``init_idx``), as the C ABI does.
"""
import numpy as np
import torch

from .tf_ops import l2_normalize, log10


def random_init_idx(rows, L, K, rng):
    """Kmeans_2.py:61-63: np.random.choice(range(l), size=K, replace=False) per row."""
    return np.stack([rng.choice(L, size=K, replace=False) for _ in range(rows)]).astype(np.int32)


class KMeans:
    def __init__(self, nb_clusters, nb_tries=10, nb_iterations=10, normalize_input=True,
                 beta=None, threshold=2.5, assign_at_end=True):
        self.K = nb_clusters
        self.tries = nb_tries
        self.iters = nb_iterations
        self.normalize_input = normalize_input
        self.beta = beta
        self.threshold = threshold
        self.assign_at_end = assign_at_end

    # Kmeans_2.py:169-188
    def get_labels(self, X, centroids, notsilent):
        d2 = (((X.unsqueeze(2) - centroids.unsqueeze(1)) ** 2) * notsilent.unsqueeze(2)).sum(3)
        if self.beta is not None:
            e = torch.exp(-1.0 * self.beta * d2)
            return e / e.sum(-1, keepdim=True)
        return torch.argmin(torch.sqrt(d2), dim=2).to(torch.int32)

    # Kmeans_2.py:145-167
    def body(self, X, labels, notsilent):
        Bt, L, E = X.shape
        Xm = X * notsilent
        if self.beta is not None:
            num = (Xm.unsqueeze(2) * labels.unsqueeze(3)).sum(1)
            return num / labels.sum(1).unsqueeze(-1)
        onehot = torch.nn.functional.one_hot(labels.long(), self.K).to(X.dtype)   # [Bt,L,K]
        total = onehot.transpose(1, 2) @ Xm                                       # segment_sum
        count = onehot.sum(1).unsqueeze(-1)                                       # ones_like(X) segment_sum
        return total / count

    # Kmeans_2.py:114-143
    def get_inertia(self, X, centroids, notsilent):
        labels = self.get_labels(X, centroids, notsilent)
        if self.beta is not None:
            d2 = ((X.unsqueeze(2) - centroids.unsqueeze(1)) ** 2).sum(-1) * labels
            return (d2.sum(1) / labels.sum(1)).sum(-1)
        idx = labels.long().unsqueeze(-1).expand(-1, -1, X.shape[-1])
        dist = ((X - torch.gather(centroids, 1, idx)) ** 2).sum(-1)               # [Bt,L]
        onehot = torch.nn.functional.one_hot(labels.long(), self.K).to(X.dtype)
        total = (onehot * dist.unsqueeze(-1)).sum(1)
        count = onehot.sum(1)
        return (total / count).sum(-1)

    # Kmeans_2.py:14-111
    def fit(self, X_in, init_idx, latent=None):
        """X_in [b,L,E]; init_idx int [b*tries, K] rows of X; latent [b,L] or None.
        Returns (centroids [b,K,E], labels [b,L] int32 or [b,L,K] float)."""
        X_in = torch.as_tensor(X_in)
        b, L, E = X_in.shape
        x = l2_normalize(X_in, -1) if self.normalize_input else X_in
        X = x.unsqueeze(1).repeat(1, self.tries, 1, 1).reshape(b * self.tries, L, E)
        Bt = b * self.tries
        idx = torch.as_tensor(np.asarray(init_idx)).long()
        centroids = torch.gather(X, 1, idx.unsqueeze(-1).expand(-1, -1, E))
        if latent is not None:
            lst = torch.as_tensor(latent).reshape(b, L)
            log_lst = log10(lst.max(-1, keepdim=True).values / lst)
            ns = (log_lst < self.threshold).to(X.dtype).reshape(b, L, 1)
            notsilent = ns.repeat(self.tries, 1, 1)                               # try-major tile (:80)
        else:
            notsilent = torch.ones(Bt, L, 1, dtype=X.dtype)
        labels = self.get_labels(X, centroids, notsilent)
        for _ in range(self.iters):
            centroids = self.body(X, labels, notsilent)
            labels = self.get_labels(X, centroids, notsilent)
        inertia = self.get_inertia(X, centroids, notsilent).reshape(b, self.tries)
        # tf.argmin (Kmeans_2.py:99) lowers to Eigen's ArgMin reducer: `if (v < best) best = v`
        # starting from +max, so a NaN inertia (empty cluster -> 0/0) is never selected and the
        # first minimum wins ties; an all-NaN row yields index 0.
        bests = torch.argmin(torch.where(torch.isnan(inertia), torch.full_like(inertia, float("inf")), inertia), 1)
        index = bests + torch.arange(b) * self.tries
        centroids = centroids[index]
        if self.assign_at_end:
            labels = self.get_labels(x, centroids, torch.ones(b, L, 1, dtype=X.dtype))
        else:
            labels = labels[index]
        self.last_inertia = inertia
        self.last_best = bests
        return centroids, labels
