"""Generates the fixtures under tests/golden/ from the read-only reference tree (run in the build
container, where /root/reference exists; the GPU box only sees the committed outputs).

    python tests/golden/make_golden.py

Outputs
  ckpt_index.json          every variable (key, shape) of the 6 shipped agents' 3 checkpoints
                           (weights/stage-*/{dynamics_model,policy_net,value_net}.index) + totals
  ckpt_s5_curriculum.npz   the trained weights of weights/stage-s5-curriculum, keyed by oracle/spec.py
                           names, stored as float16 (realistic parameter / BN moving-statistic fixture;
                           values are rounded, they are inputs of parity tests, not expected outputs)
  known_answers.json       closed-form answers quoted by the reference source itself
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ckpt, spec  # noqa: E402

REF = '/root/reference'


def main():
    stages = sorted(d for d in os.listdir(f'{REF}/weights') if os.path.isdir(f'{REF}/weights/{d}'))
    index = {}
    for s in stages:
        summ = ckpt.index_summary(f'{REF}/weights/{s}')
        totals = {m: int(sum(np.prod(shape) for _, shape in v)) for m, v in summ.items()}
        index[s] = dict(variables=summ, totals=totals)
    with open(os.path.join(HERE, 'ckpt_index.json'), 'w') as f:
        json.dump(index, f, indent=0, sort_keys=True)

    dyn, pol, val = ckpt.load_reference_checkpoint(f'{REF}/weights/stage-s5-curriculum')
    flat = {}
    for prefix, d in (('dyn/', dyn), ('pol/', pol), ('val/', val)):
        for k, v in d.items():
            flat[prefix + k] = v.astype(np.float16)
    np.savez_compressed(os.path.join(HERE, 'ckpt_s5_curriculum.npz'), **flat)

    known = dict(
        decompose_number=[[2.34, 0.234, 1.0]],                      # rl/utils.py:141-144 docstring
        shuffle_c8=[0, 2, 4, 6, 1, 3, 5, 7],                        # core/architectures.py:115-117 for C = 8
        spatial_90x120=[[44, 59], [22, 30], [11, 15], [6, 8], [3, 4]],   # App. A.2 (valid stem, SAME pool/stride-2 units)
        totals=dict(dynamics_model=2145014, policy_net=272134, value_net=271492),
    )
    with open(os.path.join(HERE, 'known_answers.json'), 'w') as f:
        json.dump(known, f, indent=1)
    print('wrote', os.listdir(HERE))


if __name__ == '__main__':
    main()
