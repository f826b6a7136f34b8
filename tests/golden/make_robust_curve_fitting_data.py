"""Extracts the 67 (x, y) pairs of the reference's RobustCurveFitting example (its own outliers included) into a fixture.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_robust_curve_fitting_data.py
Source: examples/src/main/scala/org/somelightprojections/skeres/examples/RobustCurveFitting.scala:21-89 (data; the lines
marked "Outlier point" are :41-42), :107 cauchyLoss(0.5), :117 setMaxNumIterations(25), :118 DENSE_QR, start (0, 0) :100-103.
"""
import json, re, pathlib
path = "/root/reference/examples/src/main/scala/org/somelightprojections/skeres/examples/RobustCurveFitting.scala"
lines = pathlib.Path(path).read_text().splitlines()
first = next(i for i, l in enumerate(lines) if "val Data = Vector(" in l)
last = next(i for i, l in enumerate(lines) if "case class ExponentialResidual" in l)
nums, outliers = [], []
for i in range(first, last):
    found = re.findall(r"[-+]?\d\.\d+e[-+]\d+", lines[i])
    if len(found) == 2:
        if "Outlier" in lines[i]:
            outliers.append(len(nums) // 2)
        nums += [float(t) for t in found]
assert len(nums) == 134, len(nums)
out = {"source": "RobustCurveFitting.scala:%d-%d" % (first + 2, last - 1), "outlier_indices": outliers, "cauchy_a": 0.5, "max_num_iterations": 25,
       "x": nums[0::2], "y": nums[1::2]}
pathlib.Path(__file__).with_name("robust_curve_fitting_data.json").write_text(json.dumps(out, indent=0))
print(len(out["x"]), "pairs, outliers at", outliers, [(out["x"][i], out["y"][i]) for i in outliers])
