"""evaluate.py -- SBD, |DiC| and foreground dice over a prediction directory
(/root/reference/code/evaluate.py:6-11,60-112; the relative ../data/... paths of :61-68 are arguments here).

    python code/evaluate.py --pred_dir out_dir --dataset CVPPP --names validation_image_paths.txt \
        --counts number_of_instances.txt --gt_dir .../CVPPP2017_LSC_training/training/A1
"""
import argparse
import os

import numpy as np
from PIL import Image

import _common  # noqa: F401  (import path)
from isa_b200.metrics import calc_dic, calc_dice, calc_sbd

parser = argparse.ArgumentParser()
parser.add_argument('--pred_dir', required=True, help='Prediction directory')
parser.add_argument('--dataset', type=str, required=True, help='Name of the dataset which is "CVPPP"')
parser.add_argument('--names', required=True, help='validation_image_paths.txt')
parser.add_argument('--counts', required=True, help='number_of_instances.txt (name,count)')
parser.add_argument('--gt_dir', required=True, help='directory with <name>_label.png and <name>_fg.png')
parser.add_argument('--empty_as_zero', action='store_true', help='score an image without predicted / annotated instances SBD 0 '
                    'instead of raising like the reference (evaluate.py:31-38)')

if __name__ == '__main__':
    opt = parser.parse_args()
    assert opt.dataset in ['CVPPP', ]
    names = np.loadtxt(opt.names, dtype='str', delimiter=',', ndmin=1)
    names = np.array([os.path.splitext(os.path.basename(n))[0] for n in names])
    n_objects_gts = np.loadtxt(opt.counts, dtype='str', delimiter=',', ndmin=2)
    dics, sbds, fg_dices = [], [], []
    for name in names:
        if not os.path.isfile('{}/{}/{}-n_objects.npy'.format(opt.pred_dir, name, name)):
            continue
        base = name.replace('_rgb', '')
        n_objects_gt = int(n_objects_gts[n_objects_gts[:, 0] == base][0][1])
        n_objects_pred = np.load('{}/{}/{}-n_objects.npy'.format(opt.pred_dir, name, name))
        ins_seg_gt = np.array(Image.open(os.path.join(opt.gt_dir, base + '_label.png')))
        ins_seg_pred = np.array(Image.open(os.path.join(opt.pred_dir, name, name + '-ins_mask.png')))
        fg_seg_gt = np.array(Image.open(os.path.join(opt.gt_dir, base + '_fg.png')))
        fg_seg_pred = np.array(Image.open(os.path.join(opt.pred_dir, name, name + '-fg_mask.png')))
        fg_seg_gt = (fg_seg_gt == 1).astype('bool')
        fg_seg_pred = (fg_seg_pred == 255).astype('bool')
        if ins_seg_pred.max() == 0 or ins_seg_gt.max() == 0:
            # no predicted (or no annotated) instance: the reference's np.max([]) raises here (evaluate.py:31-38) and so
            # does this script; --empty_as_zero scores such an image 0 instead (pred_list.py --keep_going writes an empty
            # mask for an image that could not be clustered) and says so
            if not opt.empty_as_zero:
                raise ValueError('zero-size array to reduction operation maximum which has no identity '
                                 '(%s has no %s instance; --empty_as_zero scores it 0)' % (name, 'predicted' if ins_seg_pred.max() == 0 else 'annotated'))
            print('no instance in %s: SBD 0 (--empty_as_zero)' % name)
            sbds.append(0.0)
        else:
            sbds.append(calc_sbd(ins_seg_gt, ins_seg_pred))
        dics.append(calc_dic(n_objects_gt, n_objects_pred))
        fg_dices.append(calc_dice(fg_seg_gt, fg_seg_pred))
    print('MEAN SBD     : ', np.mean(sbds))
    print('MEAN |DIC|   : ', np.mean(dics))
    print('MEAN FG DICE : ', np.mean(fg_dices))
