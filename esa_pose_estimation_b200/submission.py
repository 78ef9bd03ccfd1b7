"""Result sink -- same interface as /root/reference/submission.py:6-52 (SubmissionWriter)."""
import csv
import os
from datetime import datetime


class SubmissionWriter:
    """Collects (filename, q[4], r[3]) estimates and writes the ESA submission CSV."""

    def __init__(self):
        self.test_results = []
        self.real_test_results = []

    def _append(self, filename, q, r, real):
        rec = {'filename': filename, 'q': list(q), 'r': list(r)}
        (self.real_test_results if real else self.test_results).append(rec)

    def append_test(self, filename, q, r):
        """Pose estimate of a synthetic test image (submission.py:22)."""
        self._append(filename, q, r, real=False)

    def append_real_test(self, filename, q, r):
        """Pose estimate of a real test image (submission.py:29)."""
        self._append(filename, q, r, real=True)

    def append_batch(self, filenames, pose7, real=False):
        """B200-side addition: a whole gathered [N,7] (qw,qx,qy,qz,tx,ty,tz) block at once."""
        pose7 = pose7.detach().cpu().numpy() if hasattr(pose7, "detach") else pose7
        for name, p in zip(filenames, pose7):
            self._append(name, p[:4], p[4:7], real)

    def export(self, out_dir='', suffix=None):
        """Writes `submission_<suffix>.csv` into out_dir: one `filename, q0..q3, r0..r2` row per estimate,
        synthetic-test rows first and real-test rows after them, each group ordered by file name
        (the file the ESA server expects; interface of submission.py:36-52).  Returns the path."""
        from itertools import chain
        tag = suffix if suffix is not None else datetime.now().strftime("%Y%m%d-%H%M")
        path = os.path.join(out_dir, 'submission_%s.csv' % tag)
        by_name = lambda rec: rec['filename']
        rows = ([rec['filename'], *rec['q'], *rec['r']]
                for rec in chain(sorted(self.test_results, key=by_name), sorted(self.real_test_results, key=by_name)))
        with open(path, 'w', newline='') as fh:
            csv.writer(fh, lineterminator='\n').writerows(rows)
        print('Submission saved to {}.'.format(path))
        return path
