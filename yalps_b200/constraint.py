"""Constraint helpers, same meaning as src/constraint.ts:7,13,19,25."""
from __future__ import annotations


def less_eq(value: float) -> dict:
    """`{max: value}` (src/constraint.ts:7)."""
    return {"max": value}


def greater_eq(value: float) -> dict:
    """`{min: value}` (src/constraint.ts:13)."""
    return {"min": value}


def equal_to(value: float) -> dict:
    """`{equal: value}` (src/constraint.ts:19)."""
    return {"equal": value}


def in_range(lower: float, upper: float) -> dict:
    """`{min: lower, max: upper}` (src/constraint.ts:25)."""
    return {"min": lower, "max": upper}


# the reference's spellings
lessEq, greaterEq, equalTo, inRange = less_eq, greater_eq, equal_to, in_range
