"""proFit include file for the B200 backend.

proFit executes every path listed under ``include:`` in ``profit.yaml`` as a stand-alone module
(``profit/util/util.py:64-82``), so a file inside a package cannot be included directly (its relative imports would
fail).  Point ``include:`` at this file instead; it puts the repository on ``sys.path`` and imports the package
modules that register ``Surrogate["B200"]``, ``Surrogate["B200MultiOutputGP"]`` and the ``b200_*`` acquisition
functions:

    include: /path/to/repo/profit_b200_include.py
    fit: {surrogate: B200}
    active_learning: {algorithm: {class: simple, acquisition_function: {class: b200_simple_exploration}}}

Set ``PROFIT_B200_TAKE_OVER=1`` to also take over the reference's own labels ("Custom" and the acquisition labels).
"""
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import profit_b200.acquisition  # noqa: E402,F401  (registers the b200_* acquisition functions)
import profit_b200.surrogate  # noqa: E402,F401  (registers Surrogate["B200"], Surrogate["B200MultiOutputGP"])

if os.environ.get("PROFIT_B200_TAKE_OVER") == "1":
    profit_b200.surrogate.register_as_custom()
    profit_b200.acquisition.register_as_default()
