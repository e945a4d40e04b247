"""Plain-text summary of a CausalImpactAnalysis.

Presentation only -- outside the hot path (SURVEY section 2 marks the reference's
summary.py / plot.py out of scope); provided so that code written as
``causalimpact.summary(impact)`` (reference summary.py:133-178) keeps working
after switching packages.  Built with plain string formatting (no Jinja); it
reads only ``impact.summary``, like the reference does (summary.py:150-176).
"""
from __future__ import annotations

from typing import Optional


def _ci_label(alpha: float) -> str:
  return ("%g" % ((1.0 - alpha) * 100.0)) + "% CI"


def _row(label: str, avg: str, cum: str) -> str:
  return f"{label:<26}{avg:<19}{cum}"


def summary(ci_model, output_format: str = "summary", alpha: Optional[float] = None) -> str:
  """Text summary ('summary') or a short prose report ('report')."""
  table = ci_model.summary
  inferred = float(table["alpha"].mean())
  if alpha is not None and alpha != inferred:
    raise DeprecationWarning(
        "Supplying an argument to `alpha` is deprecated, since it is inferred from `ci_model`. "
        f"Set `alpha=None` to use alpha={inferred:.2f}, or refit with alpha={alpha}.")
  alpha = inferred
  if output_format not in ("summary", "report"):
    raise ValueError("`format` must be either 'summary' or 'report'. Got %s" % output_format)
  if alpha <= 0.0 or alpha >= 1.0:
    raise ValueError("`alpha` must be in (0, 1). Got %s" % alpha)
  avg, cum = table.loc["average"], table.loc["cumulative"]
  p_value = float(table["p_value"].iloc[0])

  def num(v):
    return f"{round(float(v), 1)}"

  def pct(v):
    return "{0:.1%}".format(float(v))

  def interval(lo, hi, fmt):
    lo, hi = sorted((float(lo), float(hi)))
    return f"[{fmt(lo)}, {fmt(hi)}]"

  if output_format == "summary":
    ci = _ci_label(alpha)
    lines = [
        "Posterior Inference {CausalImpact}",
        _row("", "Average", "Cumulative"),
        _row("Actual", num(avg.actual), num(cum.actual)),
        _row("Prediction (s.d.)", f"{num(avg.predicted)} ({round(float(avg.predicted_sd), 2)})",
             f"{num(cum.predicted)} ({round(float(cum.predicted_sd), 2)})"),
        _row(ci, interval(avg.predicted_lower, avg.predicted_upper, num),
             interval(cum.predicted_lower, cum.predicted_upper, num)),
        "",
        _row("Absolute effect (s.d.)", f"{num(avg.abs_effect)} ({round(float(avg.abs_effect_sd), 2)})",
             f"{num(cum.abs_effect)} ({round(float(cum.abs_effect_sd), 2)})"),
        _row(ci, interval(avg.abs_effect_lower, avg.abs_effect_upper, num),
             interval(cum.abs_effect_lower, cum.abs_effect_upper, num)),
        "",
        _row("Relative effect (s.d.)", f"{pct(avg.rel_effect)} ({pct(avg.rel_effect_sd)})",
             f"{pct(cum.rel_effect)} ({pct(cum.rel_effect_sd)})"),
        _row(ci, interval(avg.rel_effect_lower, avg.rel_effect_upper, pct),
             interval(cum.rel_effect_lower, cum.rel_effect_upper, pct)),
        "",
        f"Posterior tail-area probability p: {round(p_value, 3)}",
        "Posterior probability of an effect: {0:.2%}".format(1.0 - p_value),
        "",
        'For more details run the command: summary(impact, output_format="report")',
    ]
    return "\n".join(lines)

  level = "%g%%" % ((1.0 - alpha) * 100.0)
  significant = not (float(avg.rel_effect_lower) < 0.0 < float(avg.rel_effect_upper))
  direction = "increase" if float(avg.rel_effect) > 0 else "decrease"
  verdict = ("The interval excludes zero, so the effect is statistically significant at this level."
             if significant else
             "The interval includes zero, so the effect is not statistically significant at this "
             "level.")
  return "\n".join([
      "Analysis report {CausalImpact}",
      "",
      f"Post-intervention average of the response: {num(avg.actual)}; expected without the "
      f"intervention: {num(avg.predicted)} ({level} interval "
      f"{interval(avg.predicted_lower, avg.predicted_upper, num)}).",
      f"Estimated effect per time point: {num(avg.abs_effect)} "
      f"({level} interval {interval(avg.abs_effect_lower, avg.abs_effect_upper, num)}); "
      f"summed over the post-period: {num(cum.abs_effect)} "
      f"({interval(cum.abs_effect_lower, cum.abs_effect_upper, num)}) against an observed total of "
      f"{num(cum.actual)} and an expected total of {num(cum.predicted)}.",
      f"In relative terms the response showed a {direction} of {pct(avg.rel_effect)} "
      f"({level} interval {interval(avg.rel_effect_lower, avg.rel_effect_upper, pct)}). {verdict}",
      f"Posterior tail-area probability p = {round(p_value, 3)} "
      "(probability of an effect: {0:.2%}).".format(1.0 - p_value),
  ])


def plot(ci_model, **kwargs):   # pylint: disable=unused-argument
  raise NotImplementedError(
      "plotting is outside this engine's scope; the reference's causalimpact.plot(impact) accepts "
      "the object returned by causalimpact_b200.fit_causalimpact (it reads impact.series only)")
