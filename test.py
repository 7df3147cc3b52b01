"""Entry point kept from the reference: ``python test.py`` with ``config.cfg`` in the working directory
(reference test.py:1-401): scores every test frame with the trained UNet sets and prints the frame-level AUROC."""
from vec_vad_b200.pipeline import test

if __name__ == '__main__':
    test('config.cfg')
