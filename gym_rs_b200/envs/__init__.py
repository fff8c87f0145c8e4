"""gym_rs::envs"""
