"""Extracts the 67 (x, y) pairs of the reference's CurveFitting example into a fixture.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_curve_fitting_data.py
Source: examples/src/main/scala/org/somelightprojections/skeres/examples/CurveFitting.scala:22-90
"""
import json, re, pathlib
src = pathlib.Path("/root/reference/examples/src/main/scala/org/somelightprojections/skeres/examples/CurveFitting.scala").read_text()
body = src[src.index("val Data = Vector("):src.index("case class ExponentialResidual")]
nums = [float(t) for t in re.findall(r"[-+]?\d\.\d+e[-+]\d+", body)]
assert len(nums) == 134, len(nums)
out = {"source": "CurveFitting.scala:22-90", "x": nums[0::2], "y": nums[1::2]}
pathlib.Path(__file__).with_name("curve_fitting_data.json").write_text(json.dumps(out, indent=0))
print(len(out["x"]), "pairs")
