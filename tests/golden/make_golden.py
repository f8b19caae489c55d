"""Generate tests/golden/numpy_estimator_readme.npz by importing the reference's Python
statement of the algorithm (notebooks/quick-prototyping.ipynb, cell 4 `numpy_estimator`) from
/root/reference in the BUILD CONTAINER.  The reference cannot travel to the GPU box, so the
vectors are committed; this script is the provenance.

    python tests/golden/make_golden.py

What is stored (float64 unless noted):
  observations (f32)         notebook cell 3: np.random.seed(42); normal(20, 2, 50)
  posterior_sigma_mu         numpy_estimator's posterior as returned: np.meshgrid default
                             'xy' indexing makes it posterior[sigma_i, mu_j]
  mu_range, sigma_range (f32) np.linspace ranges it returns
  printed                    the two numbers the notebook/README print (axis-swapped)
  expectations_axis_correct  E[mu], E[sigma] with the C++ axis semantics on the same posterior
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_numpy_estimator():
    nb = json.load(open(os.path.join(REF, "notebooks", "quick-prototyping.ipynb")))
    ns = {}
    exec("import numpy as np\nfrom scipy.stats import norm\n", ns)
    for cell in nb["cells"]:
        src = "".join(cell["source"])
        if cell["cell_type"] == "code" and "def numpy_estimator" in src:
            # keep only the function definition: the cell's trailing lines call and print it
            body = src.split("posterior, mu_range, sigma_range = numpy_estimator")[0]
            exec(body, ns)
            return ns["numpy_estimator"]
    raise RuntimeError("numpy_estimator not found in the reference notebook")


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout absent; goldens are already committed")
    numpy_estimator = load_numpy_estimator()
    np.random.seed(42)
    obs = np.random.normal(20, 2, 50).astype(np.float32)
    posterior, mu_range, sigma_range = numpy_estimator(obs)
    printed = np.array([np.sum(mu_range * posterior.sum(axis=1)),
                        np.sum(sigma_range * posterior.sum(axis=0))])
    correct = np.array([np.sum(mu_range * posterior.sum(axis=0)),
                        np.sum(sigma_range * posterior.sum(axis=1))])
    out = os.path.join(HERE, "numpy_estimator_readme.npz")
    np.savez_compressed(out, observations=obs, posterior_sigma_mu=posterior, mu_range=mu_range,
                        sigma_range=sigma_range, printed=printed, expectations_axis_correct=correct)
    print("wrote", out, "printed", printed, "axis-correct", correct)


if __name__ == "__main__":
    main()
