"""CPU restatement (numpy float64) of the class-balance pixel weights -- TEST INFRASTRUCTURE ONLY.

Follows datasets/Base.py:73-89 (`BaseDataSet.get_label`) of the reference, per image:

    label_balance = label with ignore_label -> num_classes
    class_num     = np.bincount(label_balance, minlength=num_classes + 1)[:-1]
    balance == 1:   weight_class = 1 / (class_num + 1)
    balance == 2:   weight_class = (1 + 1e-8 - beta ** class_num[cls]) / (1 + 1e-8 - beta ** class_num)
    weight_class  = np.clip(weight_class, 0.0, 1.0);  append 0 for the ignore bin;  weight = weight_class[label_balance]

Pinned against the unmodified reference method by tests/golden/balance.npz.
"""
import numpy as np


def class_balance_weights(label, num_classes, sample_class=None, mode=2, beta=0.9999, ignore_label=255):
    """label: [H, W] integer array of ONE image -> (weight float64 [H, W], class_num int64 [num_classes])."""
    label_balance = np.asarray(label).astype(np.int64).copy()
    label_balance[label_balance == ignore_label] = num_classes
    class_num = np.bincount(label_balance.reshape(-1), minlength=num_classes + 1)[:-1]
    if mode == 1:
        weight_class = 1 / (class_num + 1)
    else:
        weight_class = (1 + 1e-8 - beta ** class_num[sample_class]) / (1 + 1e-8 - beta ** class_num)
    weight_class = np.clip(weight_class, 0.0, 1.0)
    weight_class = np.append(weight_class, 0)
    return weight_class[label_balance], class_num
