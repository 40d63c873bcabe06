"""PCAEncoder with the reference's interface (models/encoder.py:96-160).

Fitting (``train``) is scikit-learn on the host, as in the reference; the latent space used for
routing at predict time (``get_latent_space``) is the fesr_route kernel.  The other encoders of
the reference (VAE / spectrum / DMD) expect dense grids, not graphs, and are out of scope.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from joblib import dump, load
from sklearn.decomposition import PCA

from .. import ops


class Encoder():
    def __init__(self, n_components):
        self.n_components = n_components

    def train(self, dataset):
        pass

    def get_latent_space(self, dataset):
        pass


class PCAEncoder(Encoder):
    min_length = 280          # hard-coded at models/encoder.py:152

    def __init__(self, n_components, **kwargs):
        super().__init__(n_components)
        self.model = PCA(n_components=n_components)

    @staticmethod
    def _iter(dataset):
        if hasattr(dataset, "__getitem__") and hasattr(dataset, "__len__"):
            return (dataset[i] for i in range(len(dataset)))
        return iter(dataset)

    def train(self, dataset, save_model=False, path=None):
        self._train_graph(dataset, save_model, path)

    def _train_graph(self, dataset, save_model=False, path=None):
        data_space = [d.x.cpu().detach().numpy() for d in self._iter(dataset)]
        min_length = min(d.shape[0] for d in data_space)
        if min_length < self.min_length:
            raise ValueError(f"every subdomain needs >= {self.min_length} nodes for the PCA routing features "
                             f"(models/encoder.py:152 transforms the first {self.min_length}); the shortest has {min_length}")
        if min_length != self.min_length:
            # the reference fits on the shortest subdomain (:117) but transforms the first 280 nodes
            # (:152); the two only agree when they coincide, so the fit is cut to 280 as well
            min_length = self.min_length
        print(f'Min length: {min_length}')
        data_space = np.array([d[:min_length, :].reshape(-1) for d in data_space])
        print(f'PCA input shape: {data_space.shape}')
        self.model.fit(data_space)
        if save_model:
            self._save_model(path)

    def _save_model(self, path):
        dump(self.model, os.path.join(path, 'pca_encoder.joblib'))

    def get_latent_space(self, dataset):
        """[S, n_components] numpy, computed on the device (fesr_route)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        batch = getattr(dataset, "batch", None)
        if batch is not None:
            x_dev, node_ptr = dataset.x_dev, batch.node_ptr
            if x_dev is None:          # a sample made by with_host_inputs(): the fields are on the host
                x_dev = dataset.x_host.to(dev, dtype=torch.float32)
        else:
            xs = [d.x for d in self._iter(dataset)]
            sizes = np.array([int(t.shape[0]) for t in xs])
            if sizes.min() < self.min_length:
                raise ValueError(f"every subdomain needs >= {self.min_length} nodes, got {int(sizes.min())}")
            x_dev = torch.cat(xs).to(dev, dtype=torch.float32)
            node_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).to(dev)
        print(f'Min length: {self.min_length}')
        _, latent = ops.route(x_dev, node_ptr, self.model.mean_, self.model.components_, rows=self.min_length)
        print(f'PCA input shape: {(latent.shape[0], self.min_length * x_dev.shape[1])}')
        return latent.cpu().numpy()

    def load_model(self, path):
        self.model = load(os.path.join(path, 'pca_encoder.joblib'))
