"""
Extract the golden numbers the reference ships in its executed notebooks into small JSON fixtures.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden_from_reference.py

Sources (jajcayn/pygpso v0.6.1):
  examples/0-basic-optimisation.ipynb   per-iteration log of the depth-5, budget-50 run (evaluations, highest score,
                                        highest UCB) and the final best point
  examples/1-callbacks.ipynb            PostUpdateLogging output: the GPR hyper-parameter table after each of the 14 fits
The unit-test constants (tests/test_gp_surrogate.py:163-169,264-265; tests/test_optimisation.py:22-23) are copied by
hand into tests/golden/reference_kats.json together with their file:line.
"""
import json
import os
import re

REF = "/root/reference/examples"
OUT = os.path.dirname(os.path.abspath(__file__))


def cell_text(cell):
    text = ""
    for out in cell.get("outputs", []):
        if "text" in out:
            text += "".join(out["text"])
        elif "data" in out and "text/plain" in out["data"]:
            text += "".join(out["data"]["text/plain"])
    return text


def iteration_trace(nb_path):
    nb = json.load(open(nb_path))
    text = "".join(cell_text(c) for c in nb["cells"] if c["cell_type"] == "code")
    pat = re.compile(
        r"After (\d+)th iteration: \s*number of obj\. func\. evaluations: (\d+) \s*highest score: ([-\d.e]+) \s*"
        r"highest UCB: ([-\d.e]+)"
    )
    trace = [
        {"iteration": int(m[0]), "evaluations": int(m[1]), "highest_score": float(m[2]), "highest_ucb": float(m[3])}
        for m in pat.findall(text)
    ]
    best = re.search(r"GPPoint\(normed_coord=array\(\[([-\d.e]+),\s*([-\d.e]+)\]\), score_mu=([-\d.e]+)", text)
    return trace, best


def hyperparameter_rows(nb_path):
    """Rows of the tabulate_module_summary tables: name -> value, one dict per GP update."""
    nb = json.load(open(nb_path))
    text = "".join(cell_text(c) for c in nb["cells"] if c["cell_type"] == "code")
    rows, current = [], {}
    for line in text.splitlines():
        m = re.match(r"\s*GPR\.(mean_function\.c|kernel\.variance|kernel\.lengthscale\w*|likelihood\.variance)\s+.*?\s([-\d.e+]+)\s*$", line)
        if m:
            key = m.group(1).replace("lengthscale", "lengthscales") if m.group(1).endswith("lengthscale") else m.group(1)
            current[key] = float(m.group(2))
            if len(current) == 4:
                rows.append(current)
                current = {}
    return rows


if __name__ == "__main__":
    trace, best = iteration_trace(os.path.join(REF, "0-basic-optimisation.ipynb"))
    rows = hyperparameter_rows(os.path.join(REF, "1-callbacks.ipynb"))
    golden = {
        "source": "jajcayn/pygpso examples/0-basic-optimisation.ipynb (cells' logged output), examples/1-callbacks.ipynb",
        "config": {"bounds": [[-3, 5], [-3, 3]], "exploration_depth": 5, "budget": 50, "update_cycle": 1},
        "iterations": trace,
        "best_point": {"normed_coord": [float(best[1]), float(best[2])], "score_mu": float(best[3])},
        "hyperparameters_per_update": rows,
    }
    with open(os.path.join(OUT, "notebook_trace_depth5.json"), "w") as handle:
        json.dump(golden, handle, indent=1)
    print(f"{len(trace)} iterations, {len(rows)} hyper-parameter rows")
