"""KMeansClassifier with the reference's interface (models/classifier.py:18-54).

Fitting is scikit-learn on the host, as in the reference; ``cluster`` (the predict-time call)
is the fesr_cluster kernel.  MeanShift / GMM / Wasserstein k-means are alternative routers that
no shipped config selects and are out of scope.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from joblib import dump, load
from sklearn.cluster import KMeans
from sklearn.preprocessing import StandardScaler

from .. import ops


class Classifier:
    def __init__(self, n_clusters):
        self.n_clusters = n_clusters
        self.scaler = StandardScaler()

    def train(self, data):
        pass

    def _normalize(self, data):
        return self.scaler.transform(data)

    def cluster(self, data):
        pass


class KMeansClassifier(Classifier):
    def __init__(self, n_clusters):
        super().__init__(n_clusters)
        self.model = KMeans(n_clusters=n_clusters, random_state=0, n_init='auto')

    def train(self, data, save_model=False, path=None):
        data = self.scaler.fit_transform(data)
        self.model.fit(data)
        if save_model:
            self._save_model(path)

    def _save_model(self, path):
        dump(self.model, os.path.join(path, 'kmeans_classifier.joblib'))
        dump(self.scaler, os.path.join(path, 'kmeans_scaler.joblib'))

    def cluster(self, data):
        dev = torch.device("cuda", torch.cuda.current_device())
        latent = torch.as_tensor(np.asarray(data), dtype=torch.float64).to(dev)
        labels = ops.cluster(latent, self.scaler.mean_, self.scaler.scale_, self.model.cluster_centers_)
        return labels.cpu().numpy().astype(np.int64)

    def load_model(self, path):
        self.model = load(os.path.join(path, 'kmeans_classifier.joblib'))
        self.scaler = load(os.path.join(path, 'kmeans_scaler.joblib'))
