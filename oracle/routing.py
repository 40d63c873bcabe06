"""Oracle (numpy): ALDS routing = PCA transform -> StandardScaler -> nearest k-means centroid.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Reference anchors (/root/reference):
  * models/encoder.py:143-157  PCAEncoder.get_latent_space: first ``min_length = 280``
    nodes of every subdomain's ``x`` flattened row-major -> ``PCA.transform``
  * models/classifier.py:26-27,48-50  StandardScaler.transform -> KMeans.predict
Third-party arithmetic (scikit_learn==1.6.1, requirements.txt:8), restated from the
published definitions: PCA.transform(X) = (X - mean_) @ components_.T (whiten=False);
StandardScaler.transform(Z) = (Z - mean_) / scale_; KMeans.predict = argmin_c ||z - c||^2
(first minimum on ties).
"""
from __future__ import annotations

import numpy as np

MIN_LENGTH = 280     # models/encoder.py:152 (hard-coded)


def routing_features(x_list) -> np.ndarray:
    """[S, 280*4] float32: x[:280, :].reshape(-1) per subdomain (encoder.py:153)."""
    return np.stack([np.asarray(x, dtype=np.float32)[:MIN_LENGTH, :].reshape(-1) for x in x_list])


def pca_transform(feat, mean, components):
    return (feat - mean) @ components.T


def route(feat, pca_mean, pca_components, scaler_mean, scaler_scale, centroids):
    """Returns (labels[S] int64, latent[S, n_components])."""
    z = pca_transform(feat.astype(np.float64), pca_mean.astype(np.float64), pca_components.astype(np.float64))
    zs = (z - scaler_mean) / scaler_scale
    d2 = ((zs[:, None, :] - centroids[None, :, :]) ** 2).sum(-1)
    return np.argmin(d2, axis=1).astype(np.int64), z
