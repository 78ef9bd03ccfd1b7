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
        """Rows sorted by filename, test rows before real-test rows (submission.py:36-52)."""
        sorted_test = sorted(self.test_results, key=lambda k: k['filename'])
        sorted_real_test = sorted(self.real_test_results, key=lambda k: k['filename'])
        if suffix is None:
            suffix = datetime.now().strftime("%Y%m%d-%H%M")
        submission_path = os.path.join(out_dir, 'submission_{}.csv'.format(suffix))
        with open(submission_path, 'w') as f:
            csv_writer = csv.writer(f, lineterminator='\n')
            for result in (sorted_test + sorted_real_test):
                csv_writer.writerow([result['filename'], *(result['q'] + result['r'])])
        print('Submission saved to {}.'.format(submission_path))
        return submission_path
