// camera.hpp -- pinhole camera of the reference (gpugi/camera/camera.hpp:7-44, camera.cpp:76-85).
// The GLFW fly-cam (InteractiveCamera) and the GlobalConfig wiring are out of scope.
#pragma once
#include "vec.hpp"

class Camera
{
public:
	Camera() : m_aspectRatio(1.0f), m_yfov(60.0f), m_up(0.0f, 1.0f, 0.0f) {}
	Camera(const ei::Vec3& _position, const ei::Vec3& _lookat, float _aspectRatio, float _verticalFOV = 60.0f,
	       const ei::Vec3& _up = ei::Vec3(0.0f, 1.0f, 0.0f))
		: m_position(_position), m_lookat(_lookat), m_aspectRatio(_aspectRatio), m_yfov(_verticalFOV), m_up(_up) {}

	const ei::Vec3& GetPosition() const { return m_position; }
	const ei::Vec3& GetLookAt() const { return m_lookat; }
	const ei::Vec3& GetUp() const { return m_up; }
	float GetAspectRatio() const { return m_aspectRatio; }
	float GetVFov() const { return m_yfov; }
	void SetPosition(const ei::Vec3& p) { m_position = p; }
	void SetLookAt(const ei::Vec3& l) { m_lookat = l; }
	void SetAspectRatio(float a) { m_aspectRatio = a; }
	void SetVFov(float f) { m_yfov = f; }

	/// camera.cpp:76-85: W = view direction, U/V span the image plane scaled by tan(fov/2) (and aspect).
	void ComputeCameraParams(ei::Vec3& cameraU, ei::Vec3& cameraV, ei::Vec3& cameraW) const
	{
		cameraW = ei::normalize(m_lookat - m_position);
		cameraU = ei::normalize(ei::cross(cameraW, m_up));
		cameraV = ei::normalize(ei::cross(cameraU, cameraW));
		float f = tanf(m_yfov / 2.0f * (ei::PI / 180.0f));
		cameraU *= f * m_aspectRatio;
		cameraV *= f;
	}

private:
	ei::Vec3 m_position;
	ei::Vec3 m_lookat;
	float m_aspectRatio;
	float m_yfov;
	ei::Vec3 m_up;
};
