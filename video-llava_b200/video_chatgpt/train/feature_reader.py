"""How the reference's trainer reads what scripts/save_spatio_temporal_clip_features.py wrote
(reference: video_chatgpt/train/train.py:401-405 `LazySupervisedDataset.__getitem__` and :447-452
`DataCollatorForSupervisedDataset.__call__`), kept as two small functions so that the on-disk format
of the GPU extractor can be checked against its only consumer:

    load_video_features(video_folder, video_file)   one pickle = one [100+P, 1024] float16 ndarray
    collate_video_features(instances)               torch.tensor per sample, stacked when shapes agree
"""
import pickle

import torch


def load_video_features(video_folder, video_file):
    with open(f"{video_folder}/{video_file}", "rb") as f:
        return pickle.load(f)


def collate_video_features(instances):
    """instances: list of dicts with a 'video' ndarray -> the batch entry `video_spatio_temporal_features`
    ([B, 100+P, 1024] tensor, or a list when the shapes differ)."""
    features = [torch.tensor(inst["video"]) for inst in instances]
    if all(x is not None and x.shape == features[0].shape for x in features):
        return torch.stack(features)
    return features
