"""Entry point with the reference's name (calc_optical_flow.py): optical flow of a dataset with FlowNet2 on the CUDA path."""
import sys

from vec_vad_b200.optical_flow import main

if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'UCSDped2')
