"""``get_train_state`` (reference ``internlm/data/train_state.py:6``)."""
from internevo_b200.core.context import global_context as gpc
from internevo_b200.core.trainer import TrainState


def get_train_state(dataloader):
    if gpc.config.data.type == "tokenized":
        return TrainState(gpc.config, dataloader.batch_sampler)
    raise ValueError(f"dataset type {gpc.config.data.type} is not supported")
