from typing import Any, Dict, List, Tuple  # noqa: F401  (re-exported, the reference imports them from here)

AgentID = Any
