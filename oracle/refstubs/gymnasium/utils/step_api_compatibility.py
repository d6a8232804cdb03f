import numpy as np


def convert_to_terminated_truncated_step_api(step_returns, is_vector_env=False):
    if len(step_returns) == 5:
        return step_returns
    observations, rewards, dones, infos = step_returns
    if not is_vector_env:
        truncated = infos.pop("TimeLimit.truncated", False)
        return observations, rewards, dones and not truncated, dones and truncated, infos
    raise NotImplementedError


def convert_to_done_step_api(step_returns, is_vector_env=False):
    if len(step_returns) == 4:
        return step_returns
    observations, rewards, terminated, truncated, infos = step_returns
    if not is_vector_env:
        if truncated or terminated:
            infos["TimeLimit.truncated"] = truncated and not terminated
        return observations, rewards, terminated or truncated, infos
    raise NotImplementedError


def step_api_compatibility(step_returns, output_truncation_bool=True, is_vector_env=False):
    if output_truncation_bool:
        return convert_to_terminated_truncated_step_api(step_returns, is_vector_env)
    return convert_to_done_step_api(step_returns, is_vector_env)
