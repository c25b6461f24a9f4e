"""The part of the reference's models/base_model.py that lies on the flow hot path: the ``Model`` base
(``opt`` dict, ``Tensor``, ``save_dir``: base_model.py:9-15) with ``grid_sample`` / ``resample``
(base_model.py:123-136).  Checkpoint loading, pyramids and LR schedules are orchestration outside the
path."""
import os
from abc import ABC, abstractmethod

import torch

from . import networks


class Model(torch.nn.Module, ABC):

    def __init__(self, **opt):
        super(Model, self).__init__()
        self.opt = opt
        self.Tensor = torch.cuda.FloatTensor if opt.get('gpu_ids') else torch.LongTensor
        self.save_dir = os.path.join(opt.get('checkpoints_dir', '.'), opt.get('name', 'ir2rgb'))

    @abstractmethod
    def save(self, label):
        pass

    def grid_sample(self, input1, input2):
        if self.opt.get('fp16'):
            return torch.nn.functional.grid_sample(input1.float(), input2.float(), mode='bilinear',
                                                   padding_mode='border').half()
        return torch.nn.functional.grid_sample(input1, input2, mode='bilinear', padding_mode='border')

    def resample(self, image, flow):
        return networks.resample(image, flow)
