"""Error types and argument validation with the reference's contract (models/dgdm_model.py:28-35,
utils/validation.py:139-192; pinned by the reference's testing/test_suite.py:72-105,164-191)."""


class ValidationError(Exception):
    pass


class ModelConfigurationError(Exception):
    """Raised by DGDMModel.__init__ for any invalid argument."""


class ModelInferenceError(Exception):
    """Raised by DGDMModel.forward for invalid inputs or any internal failure."""


def validate_integer(value, min_val=None, max_val=None):
    try:
        v = int(value)
    except (ValueError, TypeError):
        raise ValidationError(f"Invalid integer value: {value}")
    if min_val is not None and v < min_val:
        raise ValidationError(f"Value too small: {v} < {min_val}")
    if max_val is not None and v > max_val:
        raise ValidationError(f"Value too large: {v} > {max_val}")
    return v


def validate_numeric(value, min_val=None, max_val=None):
    try:
        v = float(value)
    except (ValueError, TypeError):
        raise ValidationError(f"Invalid numeric value: {value}")
    if min_val is not None and v < min_val:
        raise ValidationError(f"Value too small: {v} < {min_val}")
    if max_val is not None and v > max_val:
        raise ValidationError(f"Value too large: {v} > {max_val}")
    return v


def validate_enum(value, allowed):
    if value not in allowed:
        raise ValidationError(f"Invalid choice: {value}. Allowed: {allowed}")
    return value


def validate_boolean(value):
    if isinstance(value, bool):
        return value
    if isinstance(value, str):
        if value.lower() in ("true", "1", "yes", "on"):
            return True
        if value.lower() in ("false", "0", "no", "off"):
            return False
        raise ValidationError(f"Invalid boolean value: {value}")
    if isinstance(value, (int, float)):
        return bool(value)
    raise ValidationError(f"Cannot convert to boolean: {type(value).__name__}")
