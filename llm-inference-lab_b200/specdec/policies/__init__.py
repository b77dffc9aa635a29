"""reference: src/specdec/policies/__init__.py -- same exports."""
from .policies import (  # noqa: F401
    AcceptancePolicy,
    ConfidenceThresholdPolicy,
    LongestPrefixPolicy,
    TopKAgreementPolicy,
    TypicalAcceptancePolicy,
    create_policy,
)
