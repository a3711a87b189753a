"""REPL plumbing shared by the two recommender mains (RecommenderMainCommon.scala:8-58)."""
from __future__ import annotations

import re
import sys
from dataclasses import dataclass

import numpy as np

_INPUT = re.compile(r"(\d+)\s*(\d+)?")       # RecommenderMainCommon.scala:16

PROMPT_HELP = ("Enter ID of the person to be provided with recommendation and ID of the target region:\n"
               " <person ID>[ <region ID>]\n"
               "If region ID is omitted, then the home region of the person will be used.\n      ")
PROMPT = "(CTRL-C for exit) <person ID>[ <region ID>]: "


@dataclass
class RecommenderTarget:
    personId: int
    homeRegionId: int
    targetRegionId: int


def parse_input(line: str):
    m = _INPUT.fullmatch(line)
    if not m:
        raise ValueError(f"Failed to parse input: {line}")          # IllegalArgumentException, :23
    return int(m.group(1)), (int(m.group(2)) if m.group(2) is not None else None)


def calc_recommender_target(persons, person_id: int, input_region):
    ids, home = persons
    hit = np.nonzero(ids == person_id)[0]
    if len(hit) == 0:
        raise LookupError(f"Person not found: {person_id}")          # NoSuchElementException, :54
    home_region = int(home[hit[0]])
    if input_region is None:
        print("Target region ID is not provided - falling back to the person's home region")   # :36
        input_region = home_region
    return RecommenderTarget(person_id, home_region, int(input_region))


def java_name(e: Exception) -> str:
    """getClass.getSimpleName of the exception the Scala code would have thrown."""
    from .engine import NoSuchElement
    if isinstance(e, LookupError):
        return "NoSuchElementException"
    if isinstance(e, (ValueError, NoSuchElement)):
        return "IllegalArgumentException"
    return type(e).__name__


def repl(handle_query) -> None:
    print(PROMPT_HELP)
    while True:
        try:
            sys.stdout.write(PROMPT)
            sys.stdout.flush()
            line = sys.stdin.readline()
            if not line:
                return
            handle_query(line.rstrip("\n"))
        except KeyboardInterrupt:
            return
        except Exception as e:      # KnnRecommenderMain.scala:44-49: print and keep looping
            msg = e.message if hasattr(e, "message") else str(e)
            print(f"{java_name(e)}: {msg}", file=sys.stderr)


def show(table_rows, columns) -> None:
    """Minimal Dataset.show(false) look-alike."""
    rows = [[str(v) for v in r] for r in table_rows]
    widths = [max([len(c)] + [len(r[i]) for r in rows]) for i, c in enumerate(columns)]
    sep = "+" + "+".join("-" * w for w in widths) + "+"
    print(sep)
    print("|" + "|".join(c.ljust(w) for c, w in zip(columns, widths)) + "|")
    print(sep)
    for r in rows:
        print("|" + "|".join(v.ljust(w) for v, w in zip(r, widths)) + "|")
    print(sep)
    print()
