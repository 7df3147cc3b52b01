"""Entry point kept from the reference: ``python train.py`` with ``config.cfg`` in the working directory
(reference train.py:1-440).  The stages live in vec_vad_b200/pipeline.py; for N GPUs launch with
``python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 train.py``."""
from vec_vad_b200.pipeline import train

if __name__ == '__main__':
    train('config.cfg')
