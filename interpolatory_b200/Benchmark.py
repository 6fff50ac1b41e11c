"""Middlebury harness (reference: src/Benchmark.py:18-114): interpolate the middle frame of every
<dataset>/<name>/frame10.png, frame11.png pair, compare with frame10i11.png, report mean PSNR / SSIM.
Same functions, arguments, printed JSON and results.txt as the reference; PSNR / SSIM are computed on the GPU
(metrics.py) instead of skimage, and PNGs are read / written with PIL instead of imageio."""
import glob
import json
import math
import os
import time
from os import path

import numpy as np

from . import metrics
from .util import sToMMSS, getETA, signal_progress, debug_flags


def _imread(p):
    from PIL import Image
    return np.asarray(Image.open(p).convert('RGB'))


def _imwrite(p, im):
    from PIL import Image
    Image.fromarray(np.asarray(im, np.uint8)).save(p)


def default_dataset_dir():
    """$MEMCI_MIDDLEBURY_DIR, else <repo>/benchmarks/middlebury (the reference keeps it beside src/, :22-24)."""
    return os.environ.get('MEMCI_MIDDLEBURY_DIR',
                          path.join(path.dirname(path.dirname(path.abspath(__file__))), 'benchmarks', 'middlebury'))


def benchmark(interpolator, output_path=None, dataset_dir=None):
    psnr = []
    ssim = []
    basedir = dataset_dir or default_dataset_dir()
    paths = sorted(glob.glob(f'{basedir}/*/frame10i11.png'))
    if not paths:
        raise FileNotFoundError(f'no <name>/frame10i11.png under {basedir}')
    cnt_done = 0
    start = int(round(time.time()))
    for path_to_truth in paths:
        test_name = path_to_truth.split(path.sep)[-2]
        frame_1 = _imread(path_to_truth.replace('frame10i11', 'frame10'))
        frame_2 = _imread(path_to_truth.replace('frame10i11', 'frame11'))
        im_true = _imread(path_to_truth)
        im_test = np.asarray(interpolator.get_benchmark_frame(frame_1, frame_2))
        if output_path is not None:
            _imwrite(path.join(output_path, f'{test_name}.png'), im_test)
        mse, s = metrics.frame_quality(im_true, im_test)
        psnr.append(metrics.psnr_from_mse(mse))
        ssim.append(s)
        cnt_done += 1
        pct = str(math.floor(100 * float(cnt_done) / len(paths))).rjust(3, ' ')
        elapsed_seconds = int(round(time.time())) - start
        progStr = (f'PROGRESS::{pct}%::Frame {cnt_done}/{len(paths)} | Time elapsed: {sToMMSS(elapsed_seconds)} | '
                   f'Estimated Time Left: {getETA(elapsed_seconds, cnt_done, len(paths))}')
        signal_progress(progStr)
        if debug_flags['debug_benchmark_progress']:
            print(progStr)
    end = int(round(time.time()))
    if debug_flags['debug_benchmark_progress']:
        print(f'PROGRESS::100%::Completed | Time taken: {sToMMSS((end - start))}')
    res = {'PSNR': np.mean(psnr), 'SSIM': np.mean(ssim), 'Time taken': sToMMSS((end - start))}
    res_json = json.dumps(res)
    if output_path is not None:
        with open(path.join(output_path, 'results.txt'), 'w') as f:
            f.write(res_json)
    print(res_json)
    return res


def get_middle_frame(interpolator, frame_1_path, frame_2_path, output_file_path, ground_truth_path=None):
    frame_1 = _imread(frame_1_path)
    frame_2 = _imread(frame_2_path)
    im_test = np.asarray(interpolator.get_benchmark_frame(frame_1, frame_2))
    _imwrite(output_file_path, im_test)
    if ground_truth_path is not None:
        mse, s = metrics.frame_quality(_imread(ground_truth_path), im_test)
        res = {'PSNR': metrics.psnr_from_mse(mse), 'SSIM': s}
        if res['PSNR'] == math.inf:
            res['PSNR'] = 'Infinity'
        print(json.dumps(res))
        return res
    return None
