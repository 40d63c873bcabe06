"""Factories and CLI of the entry scripts, with the reference's names and flags (utils.py:19-88)."""
from __future__ import annotations

import argparse
import time

import yaml

from .dataset.GraphDataset import AnsysDataset, DuctAnalysisDataset, SyntheticDuctDataset
from .models.classifier import KMeansClassifier
from .models.encoder import PCAEncoder
from .models.model import KernelNN, TEECNet


def load_yaml(path):
    with open(path, 'r') as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def get_cur_time():
    return time.strftime('%m-%d-%H-%M', time.localtime())


def init_model(type, in_channels, out_channels, **kwargs):
    if type == 'teecnet':
        return TEECNet(in_channels, out_channels=out_channels, **kwargs)
    elif type == 'neuralop':
        return KernelNN(width=kwargs['width'], ker_width=kwargs['width'], depth=kwargs['num_layers'],
                        in_width=in_channels, out_width=out_channels)
    elif type in ('fno', 'deeponet', 'graphsage'):
        raise ValueError(f'model type {type!r} is a grid model / incompatible with the scheduler call in the '
                         'reference (SURVEY.md section 2 rows 16, 17, 19) and is not part of fesr_b200')
    raise ValueError(f'Invalid model type: {type}')


def init_dataset(name, **kwargs):
    if name == 'duct':
        return DuctAnalysisDataset(**kwargs)
    elif name == 'ansys':
        return AnsysDataset(**kwargs)
    elif name == 'synthetic':
        return SyntheticDuctDataset(**kwargs)
    elif name == 'stored':              # the reference's partitioned store (mesh_*/subdomain_*), root = .npz / .h5 file
        from .dataset.store import StoredSubdomainDataset
        return StoredSubdomainDataset(**kwargs)
    raise ValueError(f'Invalid dataset name: {name}')


def init_encoder(type, n_components, **kwargs):
    if type == 'pca':
        return PCAEncoder(n_components=n_components)
    raise ValueError(f'Invalid encoder type: {type} (fesr_b200 builds the graph encoder the configs use: pca)')


def init_classifier(type, n_clusters, **kwargs):
    if type == 'kmeans':
        return KMeansClassifier(n_clusters=n_clusters)
    raise ValueError(f'Invalid classifier type: {type} (fesr_b200 builds the default router: kmeans)')


def parse_args():
    parser = argparse.ArgumentParser(description='Run ALDS experiment')
    parser.add_argument('--dataset', type=str, default='ansys', help='Name of the dataset')
    parser.add_argument('--encoder', type=str, default='pca', help='Name of the encoder')
    parser.add_argument('--classifier', type=str, default='kmeans', help='Name of the classifier')
    parser.add_argument('--model', type=str, default='neuralop', help='Name of the model')
    parser.add_argument('--exp_name', type=str, default='ansys_neuralop', help='Name of the experiment')
    parser.add_argument('--mode', type=str, default='pred', help="Mode of the experiment: train | pred (alias predict)")
    parser.add_argument('--exp_config', type=str, default='configs/exp_config/teecnet_ansys.yaml',
                        help='Path to the experiment configuration file')
    parser.add_argument('--train_config', type=str, default='configs/train_config/teecnet.yaml',
                        help='Path to the training configuration file')
    parser.add_argument('--precision', type=str, default=None, help='fp32 | tf32 (node contraction arithmetic)')
    return parser.parse_args()
