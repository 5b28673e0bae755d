"""MVectorTrainer -- only the hot-path caller ``extract_features`` (reference: mvector/trainer.py:146-175) is provided.

``extract_features`` walks the train / enroll / trials lists (``path\\tlabel`` lines), runs the front-end on every file
exactly like ``MVectorDataset.__getitem__`` does in ``mode='extract_feature'`` (reader.py:82-107: skip files shorter than
``min_duration``, resample, dB-normalise, crop to ``max_duration`` from the start, featurize ONE utterance -> [T, F]) and
writes the reference's feature cache: ``<save_dir>/<label>/<ms timestamp>.npy`` float32 [T, F] plus a
``*_features.txt`` list that the reference's reader consumes (reader.py:76-81).  The front-end runs on the fused sm_100a
kernel.  Training / evaluation / export are outside the embedding-extraction path and raise."""
import os
import time

import numpy as np
import torch
import yaml
from loguru import logger

from .audio import AudioSegment
from .data_utils.featurizer import AudioFeaturizer
from .utils.utils import dict_to_object, print_arguments


class MVectorTrainer(object):
    def __init__(self, configs, use_gpu=True, data_augment_configs=None):
        if use_gpu:
            assert torch.cuda.is_available(), 'GPU不可用'
        else:
            raise RuntimeError('use_gpu=False: the B200-native path has no CPU implementation')
        if isinstance(configs, str):
            with open(configs, 'r', encoding='utf-8') as f:
                configs = yaml.load(f.read(), Loader=yaml.FullLoader)
            print_arguments(configs=configs)
        self.configs = dict_to_object(configs)
        self.audio_featurizer = None

    def extract_features(self, save_dir='dataset/features', max_duration=100):
        """提取特征保存文件 (trainer.py:146-175)"""
        pc = self.configs.preprocess_conf
        self.audio_featurizer = AudioFeaturizer(feature_method=pc.feature_method,
                                                use_hf_model=pc.get('use_hf_model', False),
                                                method_args=pc.get('method_args', {}))
        ds = self.configs.dataset_conf.get('dataset', {})
        min_duration = ds.get('min_duration', 0.5)
        sample_rate = ds.get('sample_rate', 16000)
        use_db, target_db = ds.get('use_dB_normalization', True), ds.get('target_dB', -20)
        for data_list in (self.configs.dataset_conf.train_list, self.configs.dataset_conf.enroll_list,
                          self.configs.dataset_conf.trials_list):
            with open(data_list, 'r', encoding='utf-8') as f:
                lines = [ln for ln in f.read().splitlines() if ln.strip()]
            save_data_list = data_list.replace('.txt', '_features.txt')
            with open(save_data_list, 'w', encoding='utf-8') as out:
                for idx in range(len(lines)):
                    # reader.py:86-88,102-106: a short / unreadable file is replaced by its successor in the list
                    j, seg = idx, None
                    for _ in range(len(lines)):
                        path, label = lines[j].split('\t')
                        try:
                            cand = AudioSegment.from_file(path)
                            if cand.duration >= min_duration:
                                seg = cand
                                break
                        except Exception as e:  # noqa: BLE001  (the reference logs and moves on, reader.py:103-106)
                            logger.error(f"[{path}]特征提取失败，错误信息：{e}")
                        j = j + 1 if j < len(lines) - 1 else 0
                    if seg is None:
                        raise RuntimeError('no usable audio file in ' + data_list)
                    if seg.sample_rate != sample_rate:
                        seg.resample(sample_rate)
                    if use_db:
                        seg.normalize(target_db=target_db)
                    if seg.duration > max_duration:
                        seg.samples = seg.samples[:int(max_duration * seg.sample_rate)]
                    feature = self.audio_featurizer(torch.from_numpy(seg.samples)).squeeze(0).cpu().numpy()
                    label = int(label)
                    stamp = int(time.time() * 1000)
                    save_path = os.path.join(save_dir, str(label), f'{stamp}.npy').replace('\\', '/')
                    while os.path.exists(save_path):      # the reference's ms timestamp collides at GPU speed
                        stamp += 1
                        save_path = os.path.join(save_dir, str(label), f'{stamp}.npy').replace('\\', '/')
                    os.makedirs(os.path.dirname(save_path), exist_ok=True)
                    np.save(save_path, feature.astype(np.float32))
                    out.write(f'{save_path}\t{label}\n')
            logger.info(f'{data_list}列表中的数据已提取特征完成，新列表为：{save_data_list}')

    def train(self, *args, **kwargs):
        raise NotImplementedError('training is outside the B200 embedding-extraction path (SURVEY.md section 2, row 7)')

    def evaluate(self, *args, **kwargs):
        raise NotImplementedError('evaluate (trainer.py:403-485) is a SURVEY.md 8(f) "next" row, not lowered yet')

    def export(self, *args, **kwargs):
        raise NotImplementedError('torch.jit export does not apply to the C-ABI path')
