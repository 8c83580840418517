"""Writes tests/golden/demo_clouds.npz: four of the reference's demo clouds (evaluate/global_eval/demo_data), prepared
exactly like Global_test_dataset does (core/datasets.py:266-274 -> get_fixednum_pcd, seed 0; see
scripts/eval_demo_retrieval.prepare): real Oxford LiDAR geometry (ground plane, range-dependent density) incl. two
clouds that the reference pads with DUPLICATED points (exact k-NN / FPS ties).  Run in the build container only:

    python tests/golden/make_demo_clouds.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "scripts"))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from eval_demo_retrieval import prepare  # noqa: E402

names, clouds, ori = prepare("/root/reference")
padded = np.nonzero(ori < 8192)[0]
pick = [0, 57, int(padded[0]), int(padded[-1])]
np.savez_compressed(os.path.join(HERE, "demo_clouds.npz"), clouds=clouds[pick], ori_num=ori[pick],
                    names=np.array([names[i] for i in pick]), index=np.array(pick))
print("wrote", [names[i] for i in pick], ori[pick])
